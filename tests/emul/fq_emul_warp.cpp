// TEST INFRASTRUCTURE ONLY.  Runs the WARP-cooperative device functions of fastquick_b200/csrc/fq_dp_warp.cuh and
// fq_dp_wave.cuh on the host: a warp is 32 fibers (ucontext) that a round-robin scheduler advances from one warp-level
// primitive to the next, so __shfl_*_sync / __ballot_sync / __syncwarp behave as on the device (all 32 lanes reach every
// primitive -- the code under test only uses them with the full mask in warp-uniform control flow).  The results are
// compared with the per-lane statements of fq_device_dp.cuh (local_align / global_align / refine_gapped / sw_core), which
// are pinned to the reference elsewhere.  The product library never contains or calls this.
#include <ucontext.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

// ---- the SIMT stand-ins the headers see --------------------------------------------------------------------------------
// Fiber switch: glibc's swapcontext makes a sigprocmask system call per switch and a shuffle costs 128 switches, so on
// x86-64 the switch is six pushes and a stack-pointer exchange; elsewhere ucontext does the job (slowly).
#if defined(__x86_64__)
extern "C" void simt_switch(void **save_sp, void *load_sp);
asm(".text\n.globl simt_switch\n.type simt_switch,@function\nsimt_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size simt_switch,.-simt_switch\n");
#define SIMT_ASM_SWITCH 1
#endif
namespace simt {
constexpr int kLanes = 32;
static int g_cur = 0;
static bool g_done[kLanes];
static long long g_slot[kLanes];
#if SIMT_ASM_SWITCH
static void *g_main_sp, *g_sp[kLanes];
inline void yield() { simt_switch(&g_sp[g_cur], g_main_sp); }
#else
static ucontext_t g_main, g_ctx[kLanes];
inline void yield() { swapcontext(&g_ctx[g_cur], &g_main); }
#endif
}  // namespace simt

#define __device__
#define __forceinline__ inline
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::yield(); }
template <class T> static inline T simt_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "");
    long long raw = 0; memcpy(&raw, &v, sizeof(T));
    simt::g_slot[simt::g_cur] = raw;
    simt::yield();
    T out = v;
    if (src_lane >= 0 && src_lane < simt::kLanes) { long long r = simt::g_slot[src_lane]; memcpy(&out, &r, sizeof(T)); }
    simt::yield();
    return out;
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { return simt_exchange(v, simt::g_cur >= d ? simt::g_cur - d : -1); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return simt_exchange(v, src & 31); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return simt_exchange(v, simt::g_cur ^ m); }
static inline unsigned __ballot_sync(unsigned, bool p) {
    simt::g_slot[simt::g_cur] = p ? 1 : 0;
    simt::yield();
    unsigned m = 0;
    for (int l = 0; l < simt::kLanes; ++l) if (simt::g_slot[l]) m |= 1u << l;
    simt::yield();
    return m;
}
static inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }

#include "../../fastquick_b200/csrc/fq_dp_warp.cuh"

using namespace fqb;

// ---- running a function on the 32 fibers ------------------------------------------------------------------------------
namespace simt {
struct Job { void (*fn)(int lane, void *arg); void *arg; };
static Job g_job;
static std::vector<char> g_stacks;
#if SIMT_ASM_SWITCH
static void trampoline() {
    const int lane = g_cur;
    g_job.fn(lane, g_job.arg);
    g_done[lane] = true;
    for (;;) simt_switch(&g_sp[lane], g_main_sp);
}
static void run_warp(void (*fn)(int, void *), void *arg) {
    const size_t kStack = 256 * 1024;
    if (g_stacks.empty()) g_stacks.resize(kStack * kLanes + 64);
    g_job.fn = fn; g_job.arg = arg;
    for (int l = 0; l < kLanes; ++l) {
        uintptr_t top = ((uintptr_t)g_stacks.data() + kStack * (l + 1)) & ~(uintptr_t)15;
        void **sp = (void **)(top - 64);                  // r15 r14 r13 r12 rbx rbp | entry | (return slot of entry)
        for (int k = 0; k < 8; ++k) sp[k] = nullptr;
        sp[6] = (void *)trampoline;
        g_sp[l] = sp;
        g_done[l] = false;
    }
    for (;;) {
        bool any = false;
        for (int l = 0; l < kLanes; ++l)
            if (!g_done[l]) { any = true; g_cur = l; simt_switch(&g_main_sp, g_sp[l]); }
        if (!any) break;
    }
}
#else
static void trampoline(int lane) {
    g_job.fn(lane, g_job.arg);
    g_done[lane] = true;
    swapcontext(&g_ctx[lane], &g_main);
}
static void run_warp(void (*fn)(int, void *), void *arg) {
    const size_t kStack = 256 * 1024;
    if (g_stacks.empty()) g_stacks.resize(kStack * kLanes);
    g_job.fn = fn; g_job.arg = arg;
    for (int l = 0; l < kLanes; ++l) {
        getcontext(&g_ctx[l]);
        g_ctx[l].uc_stack.ss_sp = g_stacks.data() + kStack * l;
        g_ctx[l].uc_stack.ss_size = kStack;
        g_ctx[l].uc_link = &g_main;
        makecontext(&g_ctx[l], (void (*)())trampoline, 1, l);
        g_done[l] = false;
    }
    for (;;) {
        bool any = false;
        for (int l = 0; l < kLanes; ++l)
            if (!g_done[l]) { any = true; g_cur = l; swapcontext(&g_main, &g_ctx[l]); }
        if (!any) break;
    }
}
#endif
}  // namespace simt

struct WarpMem {                      // what warp_scratch() hands a warp on the device
    std::vector<int32_t> sm; std::vector<uint8_t> refc, ops, gb;
    WarpDp lane(int l) {
        WarpDp w; w.sm = sm.data(); w.n_ints = (int)sm.size(); w.refc = refc.data(); w.n_refc = (int)refc.size();
        w.ops = ops.data(); w.n_ops = (int)ops.size(); w.gb = gb.data(); w.n_bytes = (int)gb.size(); w.lane = l;
        return w;
    }
};

struct LocalArgs { WarpMem *m; RefWin R; ReadSeq Q; LocalResult out[32]; };
static void local_fn(int lane, void *a) {
    LocalArgs *A = (LocalArgs *)a;
    const WarpDp w = A->m->lane(lane);
    A->out[lane] = warp_local_align(A->R, A->R.l, A->Q, A->Q.len, w);
}
struct GlobalArgs { WarpMem *m; RefWin R; ReadSeq Q; int gap_end, band; GlobalResult out[32]; bool loaded[32]; };
static void global_fn(int lane, void *a) {
    GlobalArgs *A = (GlobalArgs *)a;
    const WarpDp w = A->m->lane(lane);
    A->loaded[lane] = warp_load_ref(A->R, w);
    if (!A->loaded[lane]) return;
    A->out[lane] = warp_global_any(A->R, 0, A->R.l, A->Q, 0, A->Q.len, A->gap_end, A->band, w);
}
struct RefineArgs { WarpMem *m; int64_t l_pac; const uint8_t *pac; ReadSeq Q; uint32_t pos; int ext; uint16_t cigar[FQB_MAX_CIGAR]; uint32_t pos_out; int nc[32]; };
static void refine_fn(int lane, void *a) {
    RefineArgs *A = (RefineArgs *)a;
    const WarpDp w = A->m->lane(lane);
    uint32_t pos = A->pos;
    uint16_t cig[FQB_MAX_CIGAR] = {0};
    A->nc[lane] = warp_refine_gapped(A->l_pac, A->pac, A->Q, &pos, A->ext, cig, FQB_MAX_CIGAR, w);
    if (lane == 0) { memcpy(A->cigar, cig, sizeof(cig)); A->pos_out = pos; }
}
struct SwArgs { WarpMem *m; int64_t l_pac; const uint8_t *pac; ReadSeq Q; int64_t beg; int reglen; uint16_t cigar[kSwCigarCap]; uint32_t cnt; int64_t beg_out; int nc[32]; };
static void sw_fn(int lane, void *a) {
    SwArgs *A = (SwArgs *)a;
    const WarpDp w = A->m->lane(lane);
    WarpSwCore core{w, nullptr, 0, 0};
    int64_t beg = A->beg; uint32_t cnt = 0;
    uint16_t cig[kSwCigarCap] = {0};
    A->nc[lane] = core(A->l_pac, A->pac, A->Q, &beg, A->reglen, cig, &cnt);
    if (lane == 0) { memcpy(A->cigar, cig, sizeof(cig)); A->cnt = cnt; A->beg_out = beg; }
}

static WarpMem make_mem(int ints, int refc, int ops, int gb) {
    WarpMem m; m.sm.assign(ints, 0x5a5a5a5a); m.refc.assign(refc, 0xee); m.ops.assign(ops, 0xee); m.gb.assign(gb, 0xee);
    return m;
}

extern "C" {

// mode 0: local alignment of the read against window [beg, beg + l) of pac (warp form vs local_align)
// mode 1: banded global alignment of the same (gap_end, band as given)
// mode 2: refine_gapped at position `beg` with `ext`
// mode 3: bwa_sw_core over [beg, beg + l)
// smem_ints / ref_cap / ops_cap = the shared-memory sizes the kernel under emulation would get.
// Returns 0 when the two forms agree, a positive code naming the first difference, -1 when BOTH say "does not fit".
int emulw_check(int mode, const uint8_t *pac, long long l_pac, long long beg, int l, const uint8_t *fwd, int len, int strand, int gap_end, int band, int ext,
                int smem_ints, int ref_cap, int ops_cap, int *info) {
    WarpMem m = make_mem(smem_ints, ref_cap, ops_cap, 256 * 1024);
    std::vector<int32_t> ints(6 * 1100);
    std::vector<uint8_t> bytes(600000);
    DpScratch sc; sc.ints = ints.data(); sc.n_ints = (int)ints.size(); sc.bytes = bytes.data(); sc.n_bytes = (int)bytes.size(); sc.istride = sc.bstride = 1;
    RefWin R; R.pac = pac; R.beg = beg; R.l = l;
    ReadSeq Q; Q.fwd = fwd; Q.len = len; Q.strand = strand;
    if (mode == 0) {
        LocalArgs A; A.m = &m; A.R = R; A.Q = Q;
        simt::run_warp(local_fn, &A);
        const LocalResult ref = local_align(R, l, Q, len, sc, 0);
        for (int k = 1; k < 32; ++k) {
            const LocalResult &a = A.out[k], &b = A.out[0];
            if (a.score != b.score || a.too_big != b.too_big) return 100 + k;
            if (!a.too_big && a.score >= 1 && (a.n_ops != b.n_ops || a.start_i != b.start_i || a.start_j != b.start_j || a.end_i != b.end_i || a.end_j != b.end_j)) return 100 + k;
        }
        const LocalResult &g = A.out[0];
        if (info) { info[0] = g.score; info[1] = g.n_ops; info[2] = g.start_i; info[3] = g.end_i; info[4] = g.too_big; info[5] = ref.score; }
        if (g.too_big || ref.too_big) return (g.too_big && ref.too_big) ? -1 : (g.too_big ? -2 : 1);
        if (g.score != ref.score) return 2;
        if (ref.score < 1) return 0;                                              // nothing else is defined
        if (g.n_ops != ref.n_ops) return 3;
        if (g.start_i != ref.start_i || g.start_j != ref.start_j || g.end_i != ref.end_i || g.end_j != ref.end_j) return 4;
        for (int k = 0; k < g.n_ops; ++k) if (m.ops[k] != bytes[k]) return 5;
        return 0;
    }
    if (mode == 1) {
        GlobalArgs A; A.m = &m; A.R = R; A.Q = Q; A.gap_end = gap_end; A.band = band;
        simt::run_warp(global_fn, &A);
        const GlobalResult ref = global_align(R, 0, l, Q, 0, len, gap_end, band, sc, 0);
        if (!A.loaded[0]) return -1;
        for (int k = 1; k < 32; ++k) if (A.out[k].score != A.out[0].score || A.out[k].n_ops != A.out[0].n_ops || A.out[k].too_big != A.out[0].too_big) return 100 + k;
        const GlobalResult &g = A.out[0];
        if (info) { info[0] = g.score; info[1] = g.n_ops; info[5] = ref.score; info[4] = g.too_big; }
        if (g.too_big || ref.too_big) return (g.too_big && ref.too_big) ? -1 : (g.too_big ? -2 : 1);
        if (g.score != ref.score) return 2;
        if (g.n_ops != ref.n_ops) return 3;
        for (int k = 0; k < g.n_ops; ++k) if (m.ops[k] != bytes[k]) return 5;
        return 0;
    }
    if (mode == 2) {
        RefineArgs A; A.m = &m; A.l_pac = l_pac; A.pac = pac; A.Q = Q; A.pos = (uint32_t)beg; A.ext = ext;
        simt::run_warp(refine_fn, &A);
        uint32_t pos = (uint32_t)beg;
        uint16_t cig[FQB_MAX_CIGAR] = {0};
        const int nc = refine_gapped(l_pac, pac, Q, &pos, ext, cig, FQB_MAX_CIGAR, sc);
        for (int k = 1; k < 32; ++k) if (A.nc[k] != A.nc[0]) return 100 + k;
        if (info) { info[0] = A.nc[0]; info[5] = nc; }
        if (A.nc[0] < 0 || nc < 0) return (A.nc[0] < 0 && nc < 0) ? -1 : (A.nc[0] < 0 ? -2 : 1);
        if (A.nc[0] != nc) return 2;
        if (A.pos_out != pos) return 3;
        for (int k = 0; k < nc; ++k) if (A.cigar[k] != cig[k]) return 5;
        return 0;
    }
    if (mode == 3) {
        SwArgs A; A.m = &m; A.l_pac = l_pac; A.pac = pac; A.Q = Q; A.beg = beg; A.reglen = l;
        simt::run_warp(sw_fn, &A);
        int64_t b = beg; uint32_t cnt = 0;
        uint16_t cig[kSwCigarCap] = {0};
        const int nc = sw_core(l_pac, pac, Q, &b, l, cig, &cnt, sc);
        for (int k = 1; k < 32; ++k) if (A.nc[k] != A.nc[0]) return 100 + k;
        if (info) { info[0] = A.nc[0]; info[5] = nc; }
        if (A.nc[0] < 0 || nc < 0) return (A.nc[0] < 0 && nc < 0) ? -1 : (A.nc[0] < 0 ? -2 : 1);
        if (A.nc[0] != nc) return 2;
        if (nc > 0 && (A.beg_out != b || A.cnt != cnt)) return 3;
        for (int k = 0; k < nc; ++k) if (A.cigar[k] != cig[k]) return 5;
        return 0;
    }
    return -9;
}

}  // extern "C"
