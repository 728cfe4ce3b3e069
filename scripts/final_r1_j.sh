python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py > gpurun_out/bench_r1_j_n1.json 2> gpurun_out/bench_r1_j_n1.err; echo "bench rc=$?"; head -c 700 gpurun_out/bench_r1_j_n1.json; echo
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_j.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_j.log 2>&1; echo "ncu rc=$?"
