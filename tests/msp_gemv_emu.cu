// CPU execution of the Msp solve's batched matrix-vector kernel (csrc/msp_gemv.cuh): the same __host__ __device__
// index functions the CUDA kernel calls, run CTA by CTA and thread by thread, the warp-shuffle reduction replaced by the
// same xor tree on an array.  Test infrastructure only (tests/test_msp_gemv_emulation.py builds it with nvcc as a host
// library); nothing in the product links it.
#include <cmath>
#include <vector>
#include "../fast_solver_lippmann_schwinger_b200/csrc/msp_gemv.cuh"

using namespace lsmsp;

template <int LANES, int UNR, bool XS>
static void emu_run(const Gemv2& a, const LaunchGeo& L) {
    std::vector<cd> xs(L.smem / sizeof(cd) + 1);
    std::vector<double> SR(256 * UNR), SI(256 * UNR), TR(256 * UNR), TI(256 * UNR);
    std::vector<ThreadMap> maps(256);
    for (unsigned bid = 0; bid < L.grid; ++bid) {
        if (XS)
            for (unsigned tid = 0; tid < 256; ++tid) stage_x(a, L.g, bid, tid, xs.data());
        for (unsigned tid = 0; tid < 256; ++tid) {
            maps[tid] = map_thread<LANES, UNR>(a, L.g, bid, tid);
            double sr[UNR], si[UNR];
            accumulate<LANES, UNR, XS>(a, maps[tid], xs.data(), sr, si);
            for (int u = 0; u < UNR; ++u) { SR[tid * UNR + u] = sr[u]; SI[tid * UNR + u] = si[u]; }
        }
        for (int o = LANES / 2; o > 0; o >>= 1) {
            for (unsigned tid = 0; tid < 256; ++tid)
                for (int u = 0; u < UNR; ++u) {
                    TR[tid * UNR + u] = SR[tid * UNR + u] + SR[(tid ^ o) * UNR + u];
                    TI[tid * UNR + u] = SI[tid * UNR + u] + SI[(tid ^ o) * UNR + u];
                }
            SR.swap(TR); SI.swap(TI);
        }
        for (unsigned tid = 0; tid < 256; ++tid) {
            double sr[UNR], si[UNR];
            for (int u = 0; u < UNR; ++u) { sr[u] = SR[tid * UNR + u]; si[u] = SI[tid * UNR + u]; }
            store_rows<LANES, UNR>(a, maps[tid], sr, si);
        }
    }
}

typedef void (*emu_fn)(const Gemv2&, const LaunchGeo&);
#define EMU_U(L, U) { emu_run<L, U, false>, emu_run<L, U, true> }
#define EMU_L(L) { EMU_U(L, 1), EMU_U(L, 2), EMU_U(L, 4), EMU_U(L, 8) }
static emu_fn emu_table[6][4][2] = { EMU_L(1), EMU_L(2), EMU_L(4), EMU_L(8), EMU_L(16), EMU_L(32) };

extern "C" {

// returns the grid size used (so the test can check the geometry), or -1
int emu_msp_gemv2(const Gemv2* a, int lanes_log2, int unr_log2, int xs, unsigned* geo_out, long* smem_out) {
    if (lanes_log2 < 0 || lanes_log2 > 5 || unr_log2 < 0 || unr_log2 > 3) return -1;
    Choice ch{lanes_log2, unr_log2, xs ? 1 : 0};
    const LaunchGeo L = launch_geo(a->rows_p, a->cols_p, a->nodes, ch);
    if (geo_out) { geo_out[0] = L.g.rblocks; geo_out[1] = L.g.npc; geo_out[2] = L.g.cpn; }
    if (smem_out) *smem_out = (long)L.smem;
    emu_table[lanes_log2][unr_log2][xs ? 1 : 0](*a, L);
    return (int)L.grid;
}

void emu_default_choice(int cols_p, int xmode, int* out3) {
    const Choice ch = default_choice(cols_p, xmode);
    out3[0] = ch.lanes_log2; out3[1] = ch.unr_log2; out3[2] = ch.xs;
}

int emu_sizeof_gemv2() { return (int)sizeof(Gemv2); }

}  // extern "C"
