# gen_golden.jl - golden vectors FROM THE UNMODIFIED REFERENCE, for tests/test_reference_golden.py.
#
#   LS_REFERENCE=/path/to/Fast_solver_Lippmann_Schwinger julia julia/gen_golden.jl [h]
#
# Needs Julia >= 1.0 with the reference's packages (README.md:5-9); no GPU and no libls_cuda.so.  It repeats
# examples/example.jl:30-93 (Greengard_Vico operator, Duan-Rokhlin sparsifier, sparsifying-preconditioned GMRES) at
# h = 0.01 (n = 101) by default and writes tests/golden/ref_example_*.npy: one operator apply, one FFTconvolution,
# As*b, the CSC arrays of As and Msp, the right-hand side and both residual histories.
#
# NOT EXECUTED IN THIS REPOSITORY's build image (no Julia): until somebody runs it the parity claims rest on the
# CPU oracle ("parity unpinned", DESIGN.md section 2).  tests/test_reference_golden.py picks the files up when present.
using IterativeSolvers, SpecialFunctions, SparseArrays, Distributed, SharedArrays, LinearAlgebra, FFTW, Random

const REF = get(ENV, "LS_REFERENCE", joinpath(@__DIR__, "..", "..", "Fast_solver_Lippmann_Schwinger"))
include(joinpath(REF, "src", "SparsifyingMatrix2D.jl"))      # pulls in FastConvolution.jl and Functions.jl
include(joinpath(REF, "src", "preconditioner.jl"))

const OUT = joinpath(@__DIR__, "..", "tests", "golden")

# minimal NPY (version 1.0) writer: little-endian, Fortran order irrelevant for vectors
npy_descr(::Type{Float64}) = "<f8"
npy_descr(::Type{ComplexF64}) = "<c16"
npy_descr(::Type{Int64}) = "<i8"
function write_npy(name::String, v::Vector{T}) where {T}
    hdr = "{'descr': '$(npy_descr(T))', 'fortran_order': False, 'shape': ($(length(v)),), }"
    pad = 64 - mod(10 + length(hdr) + 1, 64)
    hdr = hdr * " "^pad * "\n"
    open(joinpath(OUT, name), "w") do io
        write(io, UInt8[0x93]); write(io, "NUMPY"); write(io, UInt8[1, 0]); write(io, UInt16(length(hdr)))
        write(io, hdr); write(io, v)
    end
end

h = length(ARGS) >= 1 ? parse(Float64, ARGS[1]) : 0.01
k = 1 / h; a = 1
x = collect(-a/2:h:a/2); y = collect(-a/2:h:a/2)
(n, m) = length(x), length(y); N = n * m
X = repeat(x, 1, m)[:]; Y = repeat(y', n, 1)[:]
(ppw, D) = referenceValsTrapRule(); D0 = D[1]
nu(x, y) = 0.3 * exp.(-40 * (x .^ 2 + y .^ 2)) .* (abs.(x) .< 0.48) .* (abs.(y) .< 0.48)
fastconv = buildFastConvolution(x, y, h, k, nu, quadRule = "Greengard_Vico")
As = buildSparseA(k, X, Y, D0, n, m)
Mapproxsp = As + k^2 * (buildSparseAG(k, X, Y, D0, n, m) * spdiagm(0 => nu(X, Y)))
precond = SparsifyingPreconditioner(Mapproxsp, As)

Random.seed!(1234)
b = randn(ComplexF64, N)
u_inc = exp.(k * im * X)
rhs = -k^2 * FFTconvolution(fastconv, nu(X, Y) .* u_inc)
u = zeros(ComplexF64, N); info = gmres!(u, fastconv, rhs, Pl = precond, log = true)
u0 = zeros(ComplexF64, N); info0 = gmres!(u0, fastconv, rhs, log = true)

write_npy("ref_example_meta.npy", Float64[n, m, h, k, real(D0), imag(D0)])
write_npy("ref_example_b.npy", b)
write_npy("ref_example_apply.npy", fastconv * b)
write_npy("ref_example_fftconv.npy", FFTconvolution(fastconv, b))
write_npy("ref_example_Asb.npy", As * b)
write_npy("ref_example_precond_b.npy", precond \ b)
for (nm, A) in (("As", As), ("Msp", Mapproxsp))
    write_npy("ref_example_$(nm)_colptr.npy", A.colptr); write_npy("ref_example_$(nm)_rowval.npy", A.rowval)
    write_npy("ref_example_$(nm)_nzval.npy", A.nzval)
end
write_npy("ref_example_rhs.npy", rhs)
write_npy("ref_example_hist_precond.npy", Vector{Float64}(info[2].data[:resnorm]))
write_npy("ref_example_hist_plain.npy", Vector{Float64}(info0[2].data[:resnorm]))
write_npy("ref_example_u_precond.npy", u)
println("wrote ref_example_*.npy to ", OUT, "  (n = ", n, ", ", length(info[2].data[:resnorm]), " / ", length(info0[2].data[:resnorm]), " iterations)")
