// ctc_lin.cuh -- the LINEAR-domain warp-specialised fused CTC kernel for sm_100a.
//
// Same decomposition as ctc_pipe.cuh (a 2-CTA cluster per utterance, the alpha CTA and the
// time/label-reversed beta CTA meet in the middle, ONE fp32 lattice goes through HBM, TMA-staged
// rings), but the lattice recursion runs on PROBABILITIES, not on log-probabilities:
//
//      x      = aB + aY_prev                       (blank cell, before the emission)
//      inner  = aY + aB + skip * aY_prev           (label cell, before the emission)
//      aB'    = y_blank * x ;  aY' = y_label * inner
//
// i.e. 5 FP32 instructions per cell pair and step and NO MUFU (the log-domain kernel spends
// 4 MUFU + ~12 FP32 per pair).  The occupancy pass needs no exp either: occupancy =
// alpha * beta~ / P with beta~ the partner's PRE-emission value, so nothing is divided by y.
//
// Range: every THREAD keeps an exact power-of-two exponent `off` for its P pairs (true value =
// a * 2^off).  Neighbour values are brought to the receiver's scale with one exact multiply; a
// thread renormalises its cells to ~2^32 every min(P, 4) steps and keeps its scale high enough
// for whatever the thread under it can hand up in the meantime (see `renorm`).  Rows are stored
// with their per-thread exponents ([blank plane][label plane][exponents]).
//
// Safety net: whatever the scaling loses (a cell flushed to zero, a clamped scale) shows up
// as missing posterior mass.  The gradient warp checks |sum_s occupancy_t(s) - 1| <= kMassTol for
// EVERY frame; an utterance that fails (or whose likelihood underflows to 0, which includes
// every infeasible utterance) is flagged in `flags` and recomputed by the log-domain kernel
// (ctc_pipe_kernel, launched right after with the same grid; clusters of unflagged utterances
// exit at once).  The linear path is therefore exact to fp32 rounding or not used at all.
//
// Warp roles (R = recursion warps, 1 for targets up to 248 labels; NC = combine groups):
//   [0, R)          REC   the lattice recursion and nothing else: one step = emission loads, one
//                         shuffle pair, 5 FP32 per pair, one row out -- to HBM (first half of the
//                         sweep, for the partner CTA) or to a shared-memory ring (second half)
//   [R, (1+NC)R)    COMB  second half only: combines REC's rows with the partner's stored rows (one
//                         TMA bulk copy per chunk, 3-slot ring) into occupancies; group g takes the
//                         rows r == g (mod NC) of a chunk, two rows in flight per pass when registers
//                         allow; label cells are ADDED to their class slot as Q1.31 fixed point with
//                         native shared-memory integer atomics (order-independent, hence
//                         bit-reproducible); also the likelihood from the first combined row
//   then nA warps   SOFT  logits staging (cp.async groups) and the fused softmax, all frames of a chunk
//                         at once (a group of lanes per frame)
//   then nB warps   GRAD  gradient rows = gscale * (softmax - occupancy), posterior-mass check,
//                         zero fill of rows t >= T_b   (H = 1: ONE helper warp is SOFT and GRAD)
// One CTA barrier per chunk of TC frames hands the rings over: in iteration `it` SOFT works on
// chunk it, REC on chunk it-1, COMB on chunk it-2, GRAD on chunk it-3.  C2 (B=256, T=1000, V=48):
// 4 warps per CTA (REC, 2 x COMB, helper), 4 CTAs per SM, 128 registers; this shape class (V = 48, one
// helper, two combine groups, chunks of 4 frames) has its own instantiation with all of these as
// compile-time constants (FIX).
//
// FIX runs its FULL chunks in steady-state loops of their own, one per role and half (rec_iter / comb_iter /
// help_iter keep the general iteration for short chunks, the phase break and the drain): shared memory by
// explicit 32-bit addresses off one register-resident base, REC loads the emissions of step r + 1 before it
// stores the row of step r and publishes its rows in the shared-memory ring in BOTH halves (the combine
// warps, idle in the first half, copy them to the lattice in HBM), the helper interleaves the gradient and
// softmax chains and pulls the logits into L2 six chunks ahead.  Why: at the 128-register cap the general loop
// re-derived shared-memory bases from S2R SR_CgaCtaId, kernel parameters from the constant bank and the chunk
// geometry in every iteration (39 % of the combine warp's active samples; profiles/r02_ncu_summary.md).
//
// Other instantiations prune at compile time what their shape class never runs (the kernels are
// instruction-fetch bound as soon as the hot path is scattered over the whole template):
//   WIDE (V > 256, aligned rows; C4): two SOFT and two GRAD warps, a warp per frame, chunks of 2 frames, logits
//        rows by ONE TMA bulk copy per row (requested by the first combine warp), rows held in registers;
//   MID  (61 ... 256 classes, the reference's V = 177 included): four helpers, every softmax warp copies and waits
//        for its own frames, rows that are not 16-byte aligned in HBM arrive in whole 16-byte segments; softmax /
//        gradient passes branch-free by explicit shared addresses;
//   MID with launch bounds of 512 threads (launches of at most one CTA per SM: the reference's batches of 32 / 64):
//        15 warps -- four softmax + four gradient warps (a warp per frame) and four COPY warps that request the logits
//        rows, wait for them and publish them through the chunk barrier; the recursion / combine warps run the
//        steady-state loops of the headline class (RCL);
//   QUEUE (FIX only): the loop over a device-side utterance queue (persistent launch).
//
// Which utterance a cluster works on: (blockIdx.x / 2 + utt_rot) mod n_utt.  The host rotates the
// (length-sorted) batch by whole launch layers so that the longest utterances land on the SM pairs that
// end up with the fewest co-resident CTAs (ctc_abi.cu).
//
// Alignment trick: the alpha CTA shifts its lattice by delta = (P-1-S) mod P slots, so that the
// P partner cells a thread needs are exactly ONE partner thread's P cells, reversed: 128-bit
// conflict-free shared-memory loads and a single partner exponent per thread.
#pragma once
#include <type_traits>

#include "ctc_pipe.cuh"

namespace ctcb200 {

#ifndef CTC_LIN_TGT
#define CTC_LIN_TGT 32
#endif
#ifndef CTC_LIN_INMAX
#define CTC_LIN_INMAX 96
#endif
constexpr int kLinTarget = CTC_LIN_TGT; // a thread's largest cell is renormalised to ~2^32
constexpr int kLinInMax = CTC_LIN_INMAX; // what the neighbour hands up stays below 2^96 in my scale
constexpr int kLinFresh = -(1 << 24);   // exponent of a thread that has not received anything yet
constexpr int kLinHmax = 127 - (kLinInMax + 7) - 4;   // largest exponent applied to the partner cell alone (p * 2^h finite)
// smallest exponent applied to the partner cell alone; a more negative h (both sides far ABOVE 1 in their
// threads' scales: a thread that has just received its first value sits near 2^kLinInMax until its next
// renormalisation) puts the rest onto the product as well instead of flushing 2^h to zero
constexpr int kLinHmin = -100;
constexpr float kMassTol = 3.0e-5f;     // |sum of occupancies - 1| per frame
constexpr int kLinNone = -(1 << 28);    // exponent of a term that is exactly zero
constexpr float kQ31 = 2147483648.0f;   // label occupancies are accumulated as Q1.31 fixed point
#ifndef CTC_LIN_YD
#define CTC_LIN_YD 1
#endif
#ifndef CTC_LIN_TMA_Y
#define CTC_LIN_TMA_Y 0   // 1: logits rows by TMA bulk copies for every vocabulary (measured slower for 192-byte rows:
                          // C2 0.310 vs 0.280 ms); 0: cp.async, and TMA only for rows wider than 1 KB
#endif
#ifndef CTC_LIN_TC4
#define CTC_LIN_TC4 1     // chunk length as a compile-time constant in the YS = 80 variants
#endif
#ifndef CTC_LIN_PD
#define CTC_LIN_PD 2
#endif
#ifndef CTC_LIN_WIDE_FAST
#define CTC_LIN_WIDE_FAST 1   // WIDE: branch-free softmax / gradient passes by explicit shared addresses (see softmax_chunk)
#endif
#ifndef CTC_LIN_RCL_Y80
#define CTC_LIN_RCL_Y80 1   // <8,1,80,128,4> (aligned V <= 60 other than 48): recursion / combine warps on the steady-state loops
#endif
#ifndef CTC_LIN_PDL_EARLY
#define CTC_LIN_PDL_EARLY 1
#endif
#ifndef CTC_LIN_REDUX
#define CTC_LIN_REDUX 1    // MID with a warp per frame: row maximum by redux.sync.max.f32
#endif
#ifndef CTC_LIN_P1RING
#define CTC_LIN_P1RING 1  // FIX, first half: 1 = rows go through the shared-memory ring and the combine warps copy them to HBM;
                          // 0 = the recursion warp stores them to HBM itself
#endif
constexpr int kLinYDist = CTC_LIN_YD;   // logits are requested kLinYDist + 1 chunks before their softmax
constexpr int kLinPDist = CTC_LIN_PD;   // partner rows are requested kLinPDist chunks before COMB needs them

__device__ __forceinline__ int clamp_exp(int e) { return max(min(e, 127), -127); }
// 2^e for e in [-126, 127]; 0 for e <= -127 (flush); 2^127 above
__device__ __forceinline__ float pow2c(int e) { return __int_as_float((clamp_exp(e) + 127) << 23); }
// exponent of x >= 0 (zero / denormals give -127)
__device__ __forceinline__ int expo(float x) { return (__float_as_int(x) >> 23) - 127; }
__device__ __forceinline__ int warp_max_i(int x) { return __reduce_max_sync(0xffffffffu, x); }

// Shared-memory accesses by explicit 32-bit address (the FIX hot loops).  In a kernel with cluster dimensions
// the compiler derives every generic shared pointer from S2R SR_CgaCtaId (~50 cycles) and, at the
// 128-register cap, re-derives it several times per loop iteration; a pinned 32-bit base avoids that.
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds64(unsigned a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64u(unsigned a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds32i(unsigned a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts32i(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts64(unsigned a, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts64u(unsigned a, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void reds_add_u32(unsigned a, unsigned v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// class slot `byte_off` bytes into an occupancy row
__device__ __forceinline__ unsigned* occ_slot(unsigned* row, int byte_off) {
    return reinterpret_cast<unsigned*>(reinterpret_cast<char*>(row) + byte_off);
}
__device__ __forceinline__ void cp_async16_a(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// 16-byte cp.async of which only the first `bytes` (<= 16) are read from global memory; the rest is zero-filled
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes) : "memory");
}
// predicated stores (MID passes: a lane's last class may lie beyond V; a branch per element would diverge)
__device__ __forceinline__ void sts32_if(unsigned a, float v, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.shared.f32 [%0], %1;\n\t}" ::"r"(a), "f"(v), "r"((int)on) : "memory");
}
__device__ __forceinline__ void sts32i_if(unsigned a, int v, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.shared.s32 [%0], %1;\n\t}" ::"r"(a), "r"(v), "r"((int)on) : "memory");
}
__device__ __forceinline__ void stg32_if(float* p, float v, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"((int)on) : "memory");
}
__device__ __forceinline__ uint4 lds128u(unsigned a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
// (lanes with `on` clear keep what the registers held: initialise them)
__device__ __forceinline__ void lds128u_if(unsigned a, uint4& v, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "r"(a), "r"((int)on) : "memory");
}
__device__ __forceinline__ void sts128_if(unsigned a, float x, float y, float z, float w, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w), "r"((int)on) : "memory");
}
__device__ __forceinline__ void sts128u_if(unsigned a, unsigned x, unsigned y, unsigned z, unsigned w, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w), "r"((int)on) : "memory");
}
__device__ __forceinline__ void stg128_if(float* p, float x, float y, float z, float w, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p st.global.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w), "r"((int)on) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill_a(unsigned dst, const void* gmem_src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}
#ifndef CTC_LIN_PF
#define CTC_LIN_PF 1      // FIX: how the logits of chunk ka + CTC_LIN_PFD reach L2 ahead of their cp.async: 0 = not at all,
                          // 1 = prefetch.global.L2 of the two lines of a row, 2 = one cp.async.bulk.prefetch.L2 per row
#endif
#ifndef CTC_LIN_PFD
#define CTC_LIN_PFD 6
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(unsigned addr, unsigned parity) {
    unsigned ok = 0;
    for (unsigned spins = 0; !ok; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (spins > (1u << 24)) __trap();
    }
}

// A plane of a lattice row holds thread t's P cells at [P * t, P * t + P) -- except for P = 8, where
// the two 128-bit halves of a thread are split: cells 0..3 at [4 t, 4 t + 4), cells 4..7 at
// [HS + 4 t, HS + 4 t + 4) with HS = half a plane.  Every 128-bit access of a warp then covers 512
// contiguous bytes: no shared-memory bank conflicts, fully coalesced global stores.
template <int P>
__device__ __forceinline__ void store_row(float* dst, const float (&v)[P], int HS) {
    if constexpr (P == 8) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + HS) = make_float4(v[4], v[5], v[6], v[7]);
    } else if constexpr (P == 4) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (P == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int k = 0; k < P; ++k) dst[k] = v[k];
    }
}
template <int P>
__device__ __forceinline__ void load_row(const float* src, float (&v)[P], int HS) {
    if constexpr (P == 8) {
        const float4 a = *reinterpret_cast<const float4*>(src);
        const float4 b = *reinterpret_cast<const float4*>(src + HS);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if constexpr (P == 4) {
        const float4 a = *reinterpret_cast<const float4*>(src);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else if constexpr (P == 2) {
        const float2 a = *reinterpret_cast<const float2*>(src);
        v[0] = a.x; v[1] = a.y;
    } else {
#pragma unroll
        for (int k = 0; k < P; ++k) v[k] = src[k];
    }
}

// class slots of an occupancy row: V + 1 (padding pairs add their zeros to slot V), and a multiple
// of 32 plus 16 so that the 4 frames of a chunk fall on alternating halves of the banks
__host__ __device__ __forceinline__ int lin_occ_classes(int V) { return (V + 1 + 15) / 32 * 32 + 16; }
// lattice row = [blank plane: NP][label plane: NP][per-thread exponents: NP / P]
__host__ __device__ __forceinline__ int lin_row_stride(int NP, int P) { return 2 * NP + (NP / P + 3) / 4 * 4; }

// Shared-memory carve-up, shared by host (size) and device (pointers).
//   y:     NL x TC emission rows (Vs floats; slot V of a row holds 0 = what padding pairs gather)
//   a:     2 x TC rows REC publishes in the second half of its sweep (row stride RS)
//   stage: NS x TC partner lattice rows (TMA)
//   occ:   2 x TC occupancy rows: per recursion warp [class sums, Q1.31: VO][blank partial sums: 32]
struct LinSmem {
    int lab, y, a, stage, occ, bnd, red, flag, bars, total;  // byte offsets
    int Vs, VO, OW, ER, NL, NS;
    __host__ __device__ static int up(int x, int a) { return (x + a - 1) / a * a; }
    __host__ __device__ LinSmem(int NP, int R, int V, int TC, int RS, int ys) {
        // (+3: a row that is not 16-byte aligned in HBM is copied in whole 16-byte segments and lands up to 3
        // floats into its ring row; the softmax moves it to the front)
        Vs = ys > 0 ? ys : up(V + 1 + ((V & 3) ? 3 : 0), 4);
        VO = ys > 0 ? 80 : lin_occ_classes(V);
        OW = VO + 32;            // + 32 blank partial sums
        ER = R * OW;
        NL = kLinYDist + 5;      // requested kLinYDist+1 chunks early .. gradient 3 chunks later
        NS = kLinPDist + 1;      // requested kLinPDist chunks before COMB needs them
        int o = 0;
        lab = o;    o += up(NP * 4, 16);
        y = o;      o += up(NL * TC * Vs * 4, 16);
        a = o;      o += up(2 * TC * RS * 4, 16);
        stage = o;  o += up(NS * TC * RS * 4, 16);
        occ = o;    o += up(2 * TC * ER * 4, 16);
        bnd = o;    o += up(2 * (R + 1) * 8, 16);
        red = o;    o += 6 * 32 * 4;   // [0,64) renorm hand-over, [64..), [128..) reductions, [100,102) E0, 1/z
        flag = o;   o += 16;
        bars = o;   o += up((NL + NS) * 8, 16);
        total = o;
    }
};

// Length-balanced placement of a one-wave launch.  The batch is sorted by length, longest first
// (dataloader.py:53), and the hardware hands cluster c of a fresh launch to SM pair c mod `pairs`, layer by
// layer.  With n = full * pairs + rem clusters the first `rem` SM pairs host full + 1 clusters and the others
// `full`.  mode 1: the pairs with the extra cluster take the SHORTEST (full + 1) * rem utterances, the others
// the longest full * (pairs - rem), and inside either group consecutive layers run in opposite directions
// (boustrophedon), so that every SM pair carries about the same number of frames.  mode 2: plain
// boustrophedon over the layers of the launch.
__device__ __forceinline__ int lin_map_utt(int c, int n, int pairs, int mode) {
    if (pairs <= 0 || n <= pairs) return c;
    const int full = n / pairs, rem = n - full * pairs;
    const int L = c / pairs, j = c - L * pairs;
    if (rem == 0 || mode == 2) {
        const int width = min(pairs, n - L * pairs);
        return L * pairs + ((L & 1) ? width - 1 - j : j);
    }
    const int m = pairs - rem;                 // SM pairs with `full` clusters
    if (j >= rem) {
        const int jj = j - rem;
        return L * m + ((L & 1) ? m - 1 - jj : jj);
    }
    return full * m + L * rem + ((L & 1) ? rem - 1 - j : j);
}

// Template parameters: P pairs per thread; RC = number of recursion warps when it is known at
// compile time (1: the common case S + P <= 32 * P, every stride becomes an immediate) or 0 for
// "run time"; YS = floats per row of the emission ring (compile time) or 0 for "run time".
// FIX: the headline shape class (V = 48, one helper warp, two combine groups, 128 threads) with all
// of these as compile-time constants.
// QUEUE: persistent launch (pp.queue; instantiated for the headline shape class only -- the loop around the
// whole utterance costs the other instantiations 6 ... 18 % when it is merely present).
// MID: vocabularies of 61 ... 256 classes served by four helper warps (rows that are not 16-byte aligned --
// the reference's own V = 177 -- included): every softmax warp copies and waits for ITS OWN frames with cp.async
// groups, rows that start off a 16-byte line are copied in whole 16-byte segments, and the softmax / gradient
// passes hold a frame's (at most 16 per lane) classes in registers.
// WIDE: more than 256 classes in 16-byte aligned rows (C4): four helper warps, a warp per frame, TMA row copies;
// everything the narrower vocabularies need is compiled out (ncu on C4: 2.3 instruction-fetch stalls per issued
// instruction in the 175 KB kernel that carries every path).
// RISS (FIX only, launches with at most two co-resident CTAs per SM): in the steady state of the second half the RECURSION
// warp requests the partner's rows (TMA).  One lane's ~25 instructions around the request sit on the combine warps' chain
// otherwise, and a launch that does not fill the SMs is bound by exactly that chain (B = 74: 0.124 -> 0.118 ms); with three
// or four CTAs per SM the same move costs 4 % (C5 on one GPU 2.76 -> 2.87 ms), and as a launch PARAMETER its mere presence in
// the two loops cost C2 2 %: hence an instantiation of its own.
// VRUN (FIX only): the headline instantiation for the narrower 16-byte aligned vocabularies (V = 4 ... 44, V % 4 = 0:
// 26 letters + space + blank, 39 phones + blank, ...): the vocabulary is a run-time value, the helper masks the classes
// from V on (they softmax to 0) and guards its gradient stores; everything else is the headline code.
template <int P, int RC, int YS, int MAXT, int MINB, bool FIX = false, bool QUEUE = false, bool MID = false,
          bool WIDE = false, bool VRUN = false, bool RISS = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MAXT, MINB)
ctc_lin_kernel(const PipeParams pp, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FusedParams& p = pp.f;
#if CTC_LIN_PDL_EARLY
    // The fallback pass and the loss reduction behind this kernel are launched with programmatic stream serialization:
    // once every CTA of this grid has started, their CTAs may take whatever SM resources fall free and wait there
    // (griddepcontrol.wait) -- the ramp of their launch disappears behind the tail of this one.
    asm volatile("griddepcontrol.launch_dependents;");
#endif
    // threads per CTA: a compile-time constant wherever the instantiation fixes the warp roles (what only other
    // CTA shapes need -- e.g. the two-rows-in-flight combine pass of the 4-warp CTAs -- is then compiled out)
    // (measured and rejected: the same for the several-recursion-warp instantiations -- C3 1.186 -> 1.220 ms)
    // MID with launch bounds of 384 threads: EIGHT helper warps (a warp per frame of a chunk, 32 lanes per frame), for
    // launches that leave every CTA an SM of its own (at most 74 utterances: the reference's batches of 32 / 64,
    // deepspeech_ctc/train.py:75-100): with one CTA per SM the helpers' chains bound the iteration, not their number
    // MIDC (launch bounds of 512 threads): plus FOUR copy warps that do nothing but request the logits rows (a warp per
    // frame), wait for them and publish them through the chunk barrier -- copy + wait were 730 of a softmax warp's 1760
    // busy cycles per chunk, and the softmax warps bound the iteration
    constexpr bool MIDC = MID && MAXT == 512;
    constexpr bool MID8 = MID && (MAXT == 384 || MIDC);
    constexpr int MG = MID8 ? 32 : 16;      // MID: lanes per frame in the softmax / gradient passes
    // the steady-state loops of the recursion and combine warps (written for FIX) also serve MID8: the same CTA shape on
    // their side (one recursion warp, two combine groups, chunks of 4 frames); only the row strides of the emission ring
    // and of the occupancy rows are run-time values there (FIX: 80 and 112 floats, immediates)
    constexpr bool RCL = FIX || MID8 || (CTC_LIN_RCL_Y80 && RC == 1 && YS == 80 && MAXT == 128 && MINB == 4);
    const int NT = (FIX || (RC == 1 && MAXT == 128)) ? 128 : (MIDC ? 480 : (MID8 ? 352 : ((WIDE || MID) ? 224 : blockDim.x))), NW = NT >> 5;
    const int R = RC > 0 ? RC : pp.R, H = FIX ? 1 : (MID8 ? 8 : ((WIDE || MID) ? 4 : pp.H)), NP = RC > 0 ? 32 * P * RC : pp.NP;
    // loop invariants the compiler would otherwise re-derive inside the role loops at the register cap
    // (S2R SR_TID / SR_CgaCtaId cost ~50 cycles each): pinned in registers in the FIX instantiation
    int lane_pin = threadIdx.x & 31;
    unsigned sbase = smem_u32(smem_raw);
    asm volatile("" : "+r"(lane_pin), "+r"(sbase));
    const int lane = lane_pin;
    const int shift = pp.rotate > 0 ? (int)((blockIdx.x / pp.rotate) * R) % NW : 0;
    const int w = ((int)(threadIdx.x >> 5) + NW - shift) % NW;   // role (virtual) warp id
    // which utterance this cluster works on: the batch is sorted by length (dataloader.py:53); utt_rot
    // moves the longest utterances to the clusters whose SMs end up with the fewest co-resident CTAs
    int rev_pin = (int)(blockIdx.x & 1);
    asm volatile("" : "+r"(rev_pin));
    const bool rev = rev_pin != 0;
    const int T = p.T, V = (FIX && !VRUN) ? 48 : p.V, blank = p.blank;
    // rows of acts / grad start on 16-byte boundaries (always, in the V <= 60 emission-ring variants);
    // otherwise (the reference's own V = 177, params.py:27) the helpers use 4-byte copies and scalar stores
    const bool al = (YS == 80 || WIDE) ? true : ((V & 3) == 0 && ((p.frame_stride | p.utt_stride) & 3) == 0);
    const Clamp clp{p.use_clamp != 0, p.clamp_lo, p.clamp_hi};
    // the V <= 60 emission-ring variants (YS = 80) are only launched with chunks of 4 frames
    const int RS = RC > 0 ? lin_row_stride(32 * P * RC, P) : p.row_stride,
              TC = ((YS == 80 && CTC_LIN_TC4) || MID) ? 4 : (WIDE ? 2 : p.chunk);   // WIDE: chunks of 2 frames, a warp per frame
    // combine groups: group g takes the rows r == g (mod NC) of a chunk
    const int NC = (FIX || WIDE || MID) ? 2 : pp.D;
    const bool is_rec = w < R, is_comb = w >= R && w < (1 + NC) * R;
    const int hw = w - (1 + NC) * R;   // helper index (>= 0 for SOFT / GRAD warps)
    const int cg = is_comb ? (w - R) / R : 0;
    // REC and COMB warps share the thread <-> lattice slot mapping
    const int tid = (is_comb ? (w - R) % R : w) * 32 + lane;

    const LinSmem lay(NP, R, V, TC, RS, YS);
    const int Vs = YS > 0 ? YS : lay.Vs, VO = YS > 0 ? 80 : lay.VO, OW = VO + 32, ER = R * OW;
    const int NL = lay.NL, NS = lay.NS;
    int* s_lab = reinterpret_cast<int*>(smem_raw + lay.lab);
    float* s_y = reinterpret_cast<float*>(smem_raw + lay.y);
    float* s_a = reinterpret_cast<float*>(smem_raw + lay.a);
    float* s_stage = reinterpret_cast<float*>(smem_raw + lay.stage);
    float* s_occ = reinterpret_cast<float*>(smem_raw + lay.occ);
    float2* s_bnd = reinterpret_cast<float2*>(smem_raw + lay.bnd);
    int* s_red = reinterpret_cast<int*>(smem_raw + lay.red);         // [6][32]
    int* s_flag = reinterpret_cast<int*>(smem_raw + lay.flag);       // [0] redo, [1] no gradient rows
    uint64_t* bar_acts = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);   // [NL]
    uint64_t* bar_part = bar_acts + NL;                                      // [NS]

    // Persistent launch (more utterances than co-resident clusters): the clusters pull utterances from an
    // atomic queue in index order -- the batch is sorted by length, longest first, so this is the
    // longest-processing-time-first schedule and the tail of the launch is filled with the SHORT utterances.
    int& s_next = s_flag[3];
    int* const queue_ = QUEUE ? pp.queue : nullptr;
    for (;;) {
    int b_local;
    if (queue_ != nullptr) {
        __syncthreads();                     // the previous utterance is done with this CTA's shared memory
        if (threadIdx.x == 0 && (blockIdx.x & 1) == 0) s_next = atomicAdd(queue_, 1);
        cluster_sync_all();                  // rank 0's ticket is visible to rank 1 (release / acquire)
        {
            unsigned remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(&s_next)));
            asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(b_local) : "r"(remote) : "memory");
        }
        cluster_sync_all();                  // both CTAs have read it before rank 0 draws the next one
        if (b_local >= pp.n_utt) {
            // the last cluster to run dry re-arms the queue for the next launch on this stream
            if (threadIdx.x == 0 && (blockIdx.x & 1) == 0 &&
                atomicAdd(queue_ + 1, 1) == (int)(gridDim.x >> 1) - 1) {
                queue_[0] = 0;
                queue_[1] = 0;
            }
            break;
        }
    } else {
        if (pp.map_mode != 0) {
            b_local = lin_map_utt((int)(blockIdx.x >> 1), (int)(gridDim.x >> 1), pp.map_pairs, pp.map_mode);
        } else {
            b_local = (int)(blockIdx.x >> 1) + pp.utt_rot;
            if (b_local >= (int)(gridDim.x >> 1)) b_local -= (int)(gridDim.x >> 1);
        }
    }
    const int b = p.utt_begin + b_local;
    int Tb = p.in_lens[b], S = p.tgt_lens[b];
    if (Tb < 0 || Tb > T || S < 0 || S > NP - P) {
        if (threadIdx.x == 0) atomicOr(p.status, kStatusBadLength);
        Tb = min(max(Tb, 0), T);
        S = min(max(S, 0), NP - P);
    }
    const int32_t* tg = p.targets + p.tgt_off[b];
    int want_grad_pin = p.grad != nullptr ? 1 : 0;
    asm volatile("" : "+r"(want_grad_pin));
    const bool want_grad = want_grad_pin != 0;
    const float gscale = p.grad_scale ? p.grad_scale[b] : 1.0f;
    const size_t frame_stride = (size_t)p.frame_stride;
    const float* acts_b = p.acts + (size_t)b * (size_t)p.utt_stride;
    float* grad_b = want_grad ? p.grad + (size_t)b * (size_t)p.utt_stride : nullptr;
    const int V4 = V >> 2, V2 = V >> 1;

    // helper roles: SOFT warps [0, nA), GRAD warps [nA, H); a single helper does both
    const int nA = H >= 2 ? H / 2 : 1, nB = H >= 2 ? H - nA : 1;
    const bool is_copy = MIDC && hw >= H;     // copy warp hw - H requests frame hw - H of every chunk
    const bool isA = hw >= 0 && (H == 1 || hw < nA), isB = hw >= 0 && (H == 1 || hw >= nA) && !is_copy;
    const int ha = is_copy ? hw - H : hw, hb = H == 1 ? 0 : hw - nA;
    // wide vocabulary: logits rows come in by TMA bulk copies (one per row) instead of cp.async
    const bool wide_rows = (CTC_LIN_TMA_Y || WIDE) ? true : (YS == 0 && nA > 1 && V > 256 && al);

    // ---- GRAD warps: mandatory zero fill of gradient rows t >= T_b (no compute) --------
    if (want_grad && isB) {
        const int nrows = T - Tb;
        const int mine = (nrows + (rev ? 0 : 1)) >> 1;  // rows Tb+rev, Tb+rev+2, ...
        float* g = grad_b + (size_t)(Tb + (rev ? 1 : 0) + 2 * hb) * frame_stride;
        const size_t ginc = 2 * (size_t)nB * frame_stride;
        if (!WIDE && !al) {
            for (int r = hb; r < mine; r += nB, g += ginc)
                for (int c = lane; c < V; c += 32) g[c] = 0.f;
        } else if (!WIDE && (YS == 80 || V4 < 32)) {
            // narrow rows: the 32 lanes of a store cover 32 / V4 rows (V = 48: 12 lanes per row otherwise)
            int r = hb, c = lane;
            while (c >= V4) { c -= V4; r += nB; g += ginc; }
            while (r < mine) {
                reinterpret_cast<float4*>(g)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                c += 32;
                while (c >= V4) { c -= V4; r += nB; g += ginc; }
            }
        } else {
            for (int r = hb; r < mine; r += nB, g += ginc)
                for (int c = lane; c < V4; c += 32) reinterpret_cast<float4*>(g)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (Tb == 0) {  // torch: empty input => 0 for an empty target, +inf otherwise
        if (threadIdx.x == 0) {
            flags[2 * b + (rev ? 1 : 0)] = 0;
            if (!rev) p.nll[b] = (S == 0 || p.zero_infinity) ? 0.0f : CUDART_INF_F;
        }
        if (queue_ != nullptr) continue;
        return;  // both CTAs of the cluster take this exit
    }

    // ---- per-utterance setup (all warps) --------------------------------------------
    // slot s of this CTA holds pair i = s - delta.  SS = S + delta == P-1 (mod P) for the alpha
    // CTA's delta; the beta CTA uses delta = 0.  My slot s <-> partner blank slot SS - s, partner
    // label slot SS - 1 - s.
    const int SS = S + ((P - 1 - S) & (P - 1));
    const int delta = rev ? 0 : SS - S;
    for (int s = threadIdx.x; s < NP; s += NT) {
        const int i = s - delta;
        int c = V;  // padding pairs gather the zero slot of the y row
        if (i >= 0 && i < S) {
            c = rev ? tg[S - 1 - i] : tg[i];
            if (c < 0 || c >= V) {
                atomicOr(p.status, kStatusBadLabel);
                c = min(max(c, 0), V - 1);
            }
        }
        s_lab[s] = c;
    }
    for (int i = threadIdx.x; i < 2 * (R + 1); i += NT) s_bnd[i] = make_float2(0.f, 0.f);
    for (int i = threadIdx.x; i < 2 * TC * ER; i += NT) s_occ[i] = 0.f;   // occupancy accumulators start at 0
    if (threadIdx.x == 0) {
        s_flag[0] = 0; s_flag[1] = 0; s_flag[2] = 0;
        for (int i = 0; i < NL; ++i) mbar_init(bar_acts + i, wide_rows ? 1 : 32);   // 32 lanes' cp.async, or one TMA producer
        for (int i = 0; i < NS; ++i) mbar_init(bar_part + i, 1);    // one TMA producer
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();

    // ---- sweep geometry (see ctc_pipe.cuh; identical band logic, pair i = slot - delta) ------
    const int C = max(Tb - S, -1);
    const int Tm = Tb >> 1;
    const int n_store = rev ? (Tb - Tm) : Tm;
    float* lat_b = p.lattice + (size_t)(b - p.utt_begin) * (size_t)p.lat_utt_stride;
    const int tsign = rev ? -1 : 1, tbase = rev ? Tb - 1 : 0;
    const int n1 = (n_store + TC - 1) / TC;
    const int n2 = want_grad ? (Tb - n_store + TC - 1) / TC : (Tb > n_store ? 1 : 0);
    const int nch = n1 + n2;
    const int n_it = nch + 3;      // SOFT chunk it, REC it-1, COMB it-2, GRAD it-3
    auto chunk_at = [&](int c, int& tt0, int& rows) {  // first sweep step / row count of chunk c
        if (c < n1) { tt0 = c * TC; rows = min(TC, n_store - tt0); }
        else { tt0 = n_store + (c - n1) * TC; rows = want_grad ? min(TC, Tb - tt0) : 1; }
    };
    // logits rows of chunk ka -> slot `slot` of the emission ring: ONE TMA bulk copy per row (V * 4 bytes,
    // contiguous in HBM), completion on the slot's mbarrier (which expects ONE arrival).  Called by one lane.
    // The ring slot was last written through the generic proxy NL - 1 CTA barriers ago (softmax of an
    // earlier chunk) and last read at least one CTA barrier ago (gradient rows); as for the partner ring,
    // no fence.proxy.async is issued per chunk: it was measured at ~1200 cycles of the issuing warp on
    // B200 (ncu, 31 % of that warp's samples).
    auto issue_logits_tma = [&](int ka, int slot) {
        int tt0, rows;
        chunk_at(ka, tt0, rows);
        float* dst = s_y + (size_t)slot * TC * Vs;
        const ptrdiff_t inc = (ptrdiff_t)tsign * (ptrdiff_t)frame_stride;
        const float* src = acts_b + (ptrdiff_t)(tbase + tsign * tt0) * (ptrdiff_t)frame_stride;
        mbar_expect_tx(bar_acts + slot, (unsigned)(rows * V) * 4u);
        for (int r = 0; r < rows; ++r)
            bulk_g2s(dst + r * Vs, src + r * inc, (unsigned)V * 4u, bar_acts + slot);
    };
    const int s0 = tid * P;          // my first slot (REC / COMB)
    const int PW = P == 8 ? 4 : P;   // floats per thread in one contiguous piece of a plane
    const int s0p = tid * PW;        // where my cells start inside a plane
    const int HS = NP / 2;           // P == 8: distance between the two halves of a thread
    const int i0 = s0 - delta;       // my first pair (may be negative: leading padding)
    const int row_step = tsign * RS;

#ifdef CTC_B200_PROFILE
    // developer instrumentation: busy cycles of CTA 0 -> workspace header (u64 at +64):
    // [0/1] REC phase 1/2, [2/3] COMB, [4/5] SOFT 0, [6/7] GRAD 0, [8] wall, [9] iterations
    unsigned long long* prof = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(p.status) + 64);
    const int prof_role = w == 0 ? 0 : (w == R ? 2 : (hw == 0 ? 4 : (hw == nA ? 6 : -1)));
    const bool prof_on = blockIdx.x == 0 && lane == 0 && prof_role >= 0;
    long long prof_t0 = clock64(), prof_t1 = prof_t0;
    const long long prof_start = prof_t0;
#ifdef CTC_B200_PROFILE_SECTIONS
#define LPROF_SEC(slot) do { if (prof_on && prof_role == 4) { const long long t_ = clock64(); atomicAdd(prof + (slot), (unsigned long long)(t_ - prof_t1)); prof_t1 = t_; } } while (0)
#else
#define LPROF_SEC(slot) do {} while (0)
#endif
#define LPROF_BEGIN() do { prof_t0 = clock64(); prof_t1 = prof_t0; } while (0)
#define LPROF_END(phase2) do { if (prof_on) atomicAdd(prof + prof_role + ((phase2) ? 1 : 0), (unsigned long long)(clock64() - prof_t0)); } while (0)
#else
#define LPROF_BEGIN() do {} while (0)
#define LPROF_SEC(slot) do {} while (0)
#define LPROF_END(phase2) do {} while (0)
#endif

    if (is_rec) {
        // =============================================================================
        // REC: lattice recursion on probabilities
        // =============================================================================
        float skf[P];                    // 1 if label cell k also takes the skip transition, else 0
        const float* yk[P];              // &y[label k] in row 0 of the current chunk of the emission ring
        int lab[P];
        float aB[P], aY[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const int i = i0 + k;
            lab[k] = s_lab[s0 + k];
            skf[k] = (i >= 1 && i < S && lab[k] != s_lab[s0 + k - 1]) ? 1.0f : 0.0f;
            aB[k] = 0.f;
            aY[k] = 0.f;
        }
        int off = kLinFresh;             // my exponent: true value = a * 2^off
        if (i0 <= 0 && i0 + P > 0) {     // the thread that owns pair 0: virtual row "-1" = 1
#pragma unroll
            for (int k = 0; k < P; ++k) if (i0 + k == 0) aB[k] = 1.0f;
            off = 0;
        }
        // thread / warp activity windows in sweep steps (widened by P pairs, see ctc_pipe.cuh)
        unsigned win_store = (i0 - P <= S) ? (unsigned)(C + 3 * P) : 0u;   // (u + P) < win
        const int iw = 32 * P * w - delta;   // first pair of my warp
        const int w_first = (iw - P <= S) ? iw - P : 0x3fffffff;
        const int w_last = C + iw + 32 * P - 1 + P;
        float2* bnd_rd = s_bnd + w;              // slot 0 is the constant (0, 0)
        float2* bnd_wr = s_bnd + (R + 1) + w + 1;
        const bool lane0 = lane == 0, lane31 = (lane == 31) && R > 1;
        const int nbar = 32 * R;
        // per-step constants live in registers (the compiler would otherwise re-derive them from the
        // kernel parameters inside the unrolled steps)
        int wg_i = want_grad ? 1 : 0, c_i = C;
        const int offd = 2 * NP + tid - s0p;  // my exponent's position relative to my blank vector
        asm volatile("" : "+r"(win_store), "+r"(wg_i), "+r"(c_i));
        const bool wg = wg_i != 0;
        int u = -i0;                     // sweep step minus my first pair: tt - i0
        int rn_par = 0;

        // Renormalisation, every min(P, 4) steps, off the dependent chain.  Every thread wants its
        // largest cell at 2^kLinTarget.  What the thread UNDER me can hand up during the next period
        // is bounded by the largest cell of its top min(P, 4) pairs (a value moves one pair per step)
        // times 3^4, so I also keep my scale high enough for that bound to stay below 2^kLinInMax: the
        // per-step exchange then needs neither a test nor a branch, and neighbouring threads are
        // otherwise free to sit at very different scales (steep lattices: peaky, blank-dominated
        // logits).  Threads whose pairs can all no longer finish are cleared; a thread nothing has
        // reached yet takes its scale from below just before the first value arrives.
        constexpr int PT = P < 4 ? P : 4;
        auto renorm = [&]() {
            float m = 0.f, mt = 0.f;
#pragma unroll
            for (int k = 0; k < P; ++k) {
                const float c = fmaxf(aB[k], aY[k]);
                m = fmaxf(m, c);
                if (k >= P - PT) mt = fmaxf(mt, c);
            }
            const bool kill = u - (P - 1) > c_i + 1;           // even my last pair is dead
            const int e_me = (m > 0.f && !kill) ? off + expo(m) : kLinFresh;
            const int e_top = (mt > 0.f && !kill) ? off + expo(mt) : kLinFresh;
            int e_in = __shfl_up_sync(0xffffffffu, e_top, 1);
            if constexpr (RC == 1) {
                if (lane0) e_in = kLinFresh;
            } else {
                int* red = s_red + (rn_par ? 32 : 0);          // double buffered
                rn_par ^= 1;
                if (lane == 31) red[w + 1] = e_top;
                if (tid == 0) red[0] = kLinFresh;
                if (R > 1) named_bar_sync(1, nbar);
                if (lane0) e_in = red[w];
            }
            const int want = max(e_me - kLinTarget, e_in + 7 - kLinInMax);
            const int noff = want < kLinFresh / 2 ? kLinFresh : want;
            int sh = noff - off;                               // cells *= 2^-sh
            const bool reset = noff == kLinFresh || off == kLinFresh || sh > 126 || kill;
            sh = max(sh, -126);
            const float f = reset ? 0.f : __int_as_float((127 - sh) << 23);
#pragma unroll
            for (int k = 0; k < P; ++k) { aB[k] *= f; aY[k] *= f; }
            off = reset ? noff : off + sh;
        };
        // one recursion step: a[t] <- a[t-1]; xs / ins = the cells BEFORE the emission (scale `off`);
        // yo = float offset of this step's row inside the chunk of the emission ring
        auto advance = [&](const float* yb_ptr, int yo, float (&xs)[P], float (&ins)[P]) {
            const float yb = fabsf(yb_ptr[yo]);   // (the sign bit carries the fused Hardtanh's backward mask)
            float v = __shfl_up_sync(0xffffffffu, aY[P - 1], 1);
            int o = __shfl_up_sync(0xffffffffu, off, 1);
            if constexpr (RC == 1) {
                if (lane0) v = 0.f;
            } else {
                if (lane0) { const float2 bq = *bnd_rd; v = bq.x; o = __float_as_int(bq.y); }
            }
            if (u > c_i) v = 0.f;        // the pair below me can no longer finish: cut the dead tail
            const float am1 = v * pow2c(o - off);
#pragma unroll
            for (int k = P - 1; k >= 0; --k) {
                const float prev = k > 0 ? aY[k > 0 ? k - 1 : 0] : am1;
                const float yl = fabsf(yk[k][yo]);
                const float x = aB[k] + prev;
                const float in = fmaf(skf[k], prev, aY[k] + aB[k]);
                xs[k] = x;
                ins[k] = in;
                aB[k] = yb * x;
                aY[k] = yl * in;
            }
            if constexpr (RC != 1) {
                if (lane31) *bnd_wr = make_float2(aY[P - 1], __int_as_float(off));
            }
        };
        auto end_step = [&]() {
            ++u;
            if constexpr (RC != 1) {
                float2* t = bnd_rd; bnd_rd = bnd_wr - 1; bnd_wr = t + 1;   // flip the double buffer
                if (R > 1) named_bar_sync(1, nbar);
            }
        };
        auto warp_on = [&](int tt) { return RC == 1 ? true : (tt >= w_first && tt <= w_last); };

        Ring ring_y(NL);                  // position of the chunk REC works on
        int a_buf = 0;
        renorm();                         // normalises the start value
        auto rec_iter = [&](int it) {
            const int k = it - 1;
            LPROF_BEGIN();
            if (k >= 0 && k < nch) {
                int tt0, rows;
                chunk_at(k, tt0, rows);
                const float* ychunk = s_y + (size_t)ring_y.slot * TC * Vs;
                const float* yb_ptr = ychunk + blank;
#pragma unroll
                for (int q = 0; q < P; ++q) yk[q] = ychunk + lab[q];
                if (k < n1) {
                    // ---- first half: pre-emission rows go to HBM for the partner ----------
                    // one running offset from the lattice base; everything else is an immediate
                    long long roff = (long long)(b - p.utt_begin) * p.lat_utt_stride +
                                     (long long)(tbase + tsign * tt0) * RS + s0p;
                    asm volatile("" : "+l"(roff));
                    float* row = p.lattice + roff;
                    const int last_r = (wg || k < n1 - 1) ? -1 : rows - 1;   // forward only: just the last row
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (r >= rows) break;
                        if (warp_on(tt0 + r)) {
                            float xs[P], ins[P];
                            advance(yb_ptr, r * Vs, xs, ins);
                            if ((wg || r == last_r) && (unsigned)(u + P) < win_store) {
                                store_row<P>(row, xs, HS);
                                store_row<P>(row + NP, ins, HS);
                                *reinterpret_cast<int*>(row + offd) = off;
                            }
                        }
                        row += row_step;
                        end_step();
                        if (P < 4 && (r + 1) % P == 0 && r + 1 < rows) renorm();
                    }
                } else {
                    // ---- second half: post-emission rows go to the shared-memory ring for COMB ----
                    float* arow = s_a + (size_t)a_buf * TC * RS + s0p;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (r >= rows) break;
                        if (warp_on(tt0 + r)) {
                            float xs[P], ins[P];
                            advance(yb_ptr, r * Vs, xs, ins);
                        }
                        // (a warp outside the band publishes its zeros / stale cells: COMB masks them)
                        store_row<P>(arow + r * RS, aB, HS);
                        store_row<P>(arow + r * RS + NP, aY, HS);
                        *reinterpret_cast<int*>(arow + r * RS + offd) = off;
                        end_step();
                        if (P < 4 && (r + 1) % P == 0 && r + 1 < rows) renorm();
                    }
                    a_buf ^= 1;
                }
                renorm();
                ring_y.advance();
            }
            LPROF_END(k >= n1);
            __syncthreads();
            if (it == n1) {  // phase break (see the COMB branch)
                cluster_sync_all();
                __syncthreads();
            }
        };
        int it = 0;
        if constexpr (RCL) {
            // The headline shape class with a gradient, full chunks, in loops of their own.  Emissions of step
            // r + 1 are loaded BEFORE the row of step r is stored (loads cannot be hoisted over shared-memory
            // stores), and the first half publishes its pre-emission rows in the shared-memory ring as well: the
            // combine warps, idle until the phase break, copy them to HBM (a recursion warp that stores to HBM
            // itself waits for every store to have read its registers before it may overwrite them).
            if (wg) {
                constexpr unsigned RSB = 544u * 4u;
                const unsigned YSB = FIX ? 80u * 4u : (unsigned)Vs * 4u, YCH = 4u * YSB;
                unsigned labo[P];
#pragma unroll
                for (int q = 0; q < P; ++q) labo[q] = (unsigned)lab[q] * 4u;
                const unsigned blo = (unsigned)blank * 4u;
                const unsigned ar0 = sbase + lay.a + (unsigned)lane * 16u;
                const unsigned exo = 2048u + (unsigned)lane * 4u - (unsigned)lane * 16u;   // my exponent, from my blank cells
                auto load_y = [&](unsigned yrow, float (&y)[P + 1]) {
#pragma unroll
                    for (int q = 0; q < P; ++q) y[q] = lds32(yrow + labo[q]);
                    y[P] = lds32(yrow + blo);
                };
                // one recursion step with the emissions in registers (see `advance`)
                auto advance_y = [&](const float (&y)[P + 1], float (&xs)[P], float (&ins)[P]) {
                    const float yb = fabsf(y[P]);
                    float v = __shfl_up_sync(0xffffffffu, aY[P - 1], 1);
                    const int o = __shfl_up_sync(0xffffffffu, off, 1);
                    if (lane0) v = 0.f;
                    if (u > c_i) v = 0.f;
                    const float am1 = v * pow2c(o - off);
#pragma unroll
                    for (int k = P - 1; k >= 0; --k) {
                        const float prev = k > 0 ? aY[k > 0 ? k - 1 : 0] : am1;
                        const float x = aB[k] + prev;
                        const float in = fmaf(skf[k], prev, aY[k] + aB[k]);
                        xs[k] = x;
                        ins[k] = in;
                        aB[k] = yb * x;
                        aY[k] = fabsf(y[k]) * in;
                    }
                };
                for (; it < n_it && it < 1; ++it) rec_iter(it);
                int pbuf = 0;
                for (; it < n1; ++it) {          // chunk it - 1 in [0, n1 - 1): first half, full chunk
                    LPROF_BEGIN();
                    const unsigned yb0 = sbase + lay.y + (unsigned)ring_y.slot * YCH;
                    const unsigned ar = ar0 + (unsigned)pbuf * (4u * RSB);
#if !CTC_LIN_P1RING
                    float* grow = p.lattice + ((long long)(b - p.utt_begin) * p.lat_utt_stride +
                                               (long long)(tbase + tsign * (it - 1) * 4) * RS + s0p);
#endif
                    float ya[P + 1], yn[P + 1];
                    load_y(yb0, ya);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (r < 3) load_y(yb0 + (unsigned)(r + 1) * YSB, yn);
                        float xs[P], ins[P];
                        advance_y(ya, xs, ins);
#if CTC_LIN_P1RING
                        const unsigned a = ar + (unsigned)r * RSB;
                        sts128(a, xs[0], xs[1], xs[2], xs[3]);
                        sts128(a + 512u, xs[4], xs[5], xs[6], xs[7]);
                        sts128(a + 1024u, ins[0], ins[1], ins[2], ins[3]);
                        sts128(a + 1536u, ins[4], ins[5], ins[6], ins[7]);
                        sts32i(a + exo, off);
#else
                        if ((unsigned)(u + P) < win_store) {
                            store_row<P>(grow, xs, HS);
                            store_row<P>(grow + NP, ins, HS);
                            *reinterpret_cast<int*>(grow + offd) = off;
                        }
                        grow += row_step;
#endif
                        ++u;
#pragma unroll
                        for (int q = 0; q <= P; ++q) ya[q] = yn[q];
                    }
                    renorm();
                    ring_y.advance();
                    pbuf ^= 1;
                    LPROF_END(false);
                    __syncthreads();
                }
                for (; it < n_it && it <= n1; ++it) rec_iter(it);     // last chunk of the first half, phase break
                unsigned ps = 0u;                // RISS: ring slot of chunk `it`, (it - n1) mod NS
                if constexpr (RISS) {
                    ps = (unsigned)(it - n1) % 3u;
                    if (lane0) fence_proxy_async();
                }
                for (; it < nch; ++it) {         // chunk it - 1 in [n1, nch - 1): second half, full chunk
                    LPROF_BEGIN();
                    if constexpr (RISS) {        // partner rows of chunk `it` (the combine warps consume them in iteration it + 2)
                        if (it >= n1 + 3 && lane0) {
                            const int tt0i = n_store + (it - n1) * 4, rowsi = min(4, Tb - tt0i);
                            const int t_lo = rev ? tbase - (tt0i + rowsi - 1) : tt0i;
                            const unsigned bytes = (unsigned)rowsi * RSB, bar = sbase + lay.bars + 8u * (unsigned)lay.NL + 8u * ps;
                            mbar_expect_tx_a(bar, bytes);
                            bulk_g2s_a(sbase + lay.stage + ps * (4u * RSB), lat_b + (ptrdiff_t)t_lo * 544, bytes, bar);
                        }
                        ps = ps == 2u ? 0u : ps + 1u;
                    }
                    const unsigned yb0 = sbase + lay.y + (unsigned)ring_y.slot * YCH;
                    const unsigned ar = ar0 + (unsigned)a_buf * (4u * RSB);
                    float ya[P + 1], yn[P + 1];
                    load_y(yb0, ya);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (r < 3) load_y(yb0 + (unsigned)(r + 1) * YSB, yn);
                        float xs[P], ins[P];
                        advance_y(ya, xs, ins);
                        const unsigned a = ar + (unsigned)r * RSB;
                        sts128(a, aB[0], aB[1], aB[2], aB[3]);
                        sts128(a + 512u, aB[4], aB[5], aB[6], aB[7]);
                        sts128(a + 1024u, aY[0], aY[1], aY[2], aY[3]);
                        sts128(a + 1536u, aY[4], aY[5], aY[6], aY[7]);
                        sts32i(a + exo, off);
                        ++u;
#pragma unroll
                        for (int q = 0; q <= P; ++q) ya[q] = yn[q];
                    }
                    renorm();
                    ring_y.advance();
                    a_buf ^= 1;
                    LPROF_END(true);
                    __syncthreads();
                }
            }
        }
        for (; it < n_it; ++it) rec_iter(it);
    } else if (is_comb) {
        // =============================================================================
        // COMB: occupancies of the second half = REC's row x the partner's stored row
        // =============================================================================
        const int wc = (w - R) % R;      // which recursion warp I shadow
        int lab[P];                      // class of my label cell k, as a BYTE offset into an occupancy row
        bool vB[P], vY[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const int i = i0 + k;
            lab[k] = s_lab[s0 + k] * 4;
            vB[k] = i >= 0 && i <= S;
            vY[k] = i >= 0 && i < S;
        }
        // partner thread of my P cells (blank k <-> its blank P-1-k; label k <-> its label P-2-k,
        // my last label <-> the last label of the thread below it)
        const int M = (SS - (P - 1)) / P;
        const int X = M - tid;
        const bool hasX = X >= 0 && X < 32 * R, hasX1 = X >= 1 && X - 1 < 32 * R;
        const unsigned win_cons = (i0 <= S && hasX) ? (unsigned)(C + P) : 0u;    // (tt - i0) < win
        const int offd = 2 * NP + tid - s0p;
        const int nbar = 32 * R;
        // FIX fast path: byte offsets that never change (row A = cg, row B = cg + 2 of a full chunk; the
        // reversed sweep walks the staged chunk backwards)
        const unsigned fx_x = hasX ? (unsigned)X * 16u : 0u;                     // partner thread's cells in a plane
        const unsigned fx_stA = (unsigned)(rev ? 3 - cg : cg) * (544u * 4u) + fx_x;
        const unsigned fx_stB = (unsigned)(rev ? 1 - cg : cg + 2) * (544u * 4u) + fx_x;
        const unsigned fx_e = 2048u + (hasX ? (unsigned)X * 4u : 0u) - fx_x;    // its exponent, from its blank cells
        const unsigned fx_ar = (unsigned)cg * (544u * 4u) + (unsigned)lane * 16u;  // my cells in REC's row cg
        const unsigned fx_o = 2048u + (unsigned)lane * 4u - (unsigned)lane * 16u;  // my exponent, from my blank cells
        const unsigned fx_bl = ((FIX ? 80u : (unsigned)VO) + (unsigned)lane) * 4u;   // my blank partial sum in an occupancy row
        int E0 = 0;                      // integer part of log2 P(labels | logits)
        float rz = 0.f;                  // 1 / mantissa sum: occupancy = a * p~ * 2^(off+o-E0) * rz
        const bool iss_part = cg == 0 && wc == 0;   // this warp also requests the partner's rows (TMA)
        auto issue_partner = [&](int kp, int slot) {
            if (lane == 0) {
                int tt0, rows;
                chunk_at(kp, tt0, rows);
                uint64_t* bar = bar_part + slot;
                const int t_lo = rev ? tbase - (tt0 + rows - 1) : tt0;
                mbar_expect_tx(bar, (unsigned)(rows * RS) * 4u);
                bulk_g2s(s_stage + (size_t)slot * TC * RS, lat_b + (ptrdiff_t)t_lo * RS,
                         (unsigned)(rows * RS) * 4u, bar);
            }
        };
        Ring ring_part(NS), iss_p(NS);
        Ring iss_y(NL);                  // WIDE: this warp also requests the logits rows (TMA), kLinYDist + 1 chunks ahead
        if constexpr (WIDE) {
            if (iss_part) {
                for (int k = 0; k <= kLinYDist; ++k) {
                    if (k < nch && lane == 0) issue_logits_tma(k, iss_y.slot);
                    iss_y.advance();
                }
            }
        }
        int a_buf = 0, o_buf = 0;
        // loop-invariant kernel parameters live in registers (each re-read from the constant bank
        // would be an exposed latency in front of a branch)
        int wgc_i = want_grad ? 1 : 0, nc_i = NC, n1_i = n1, nch_i = nch;
        asm volatile("" : "+r"(wgc_i), "+r"(nc_i), "+r"(n1_i), "+r"(nch_i));
        const bool wgc = wgc_i != 0;
        auto comb_iter = [&](int it) {
                LPROF_BEGIN();
                if constexpr (WIDE) {
                    if (iss_part) {
                        if (it + kLinYDist + 1 < nch && lane == 0) issue_logits_tma(it + kLinYDist + 1, iss_y.slot);
                        iss_y.advance();
                    }
                }
                // partner rows of chunk `it` (consumed in iteration it+2); the first two consume chunks
                // are requested at the phase break
                if (it >= n1_i + 2 && it + kLinPDist - 2 < nch_i && wgc) {
                    if (iss_part) issue_partner(it + kLinPDist - 2, iss_p.slot);
                    iss_p.advance();
                }
                const int k = it - 2;
                if (k >= n1_i && k < nch_i) {
                    int tt0, rows;
                    chunk_at(k, tt0, rows);
                    // the chunk with the first combined row is done by group 0 alone (it yields E0 and
                    // 1/z, which the other groups pick up from shared memory one barrier later)
                    const bool first = k == n1_i;
                    if (k == n1_i + 1 && cg > 0) { E0 = s_red[100]; rz = __int_as_float(s_red[101]); }
                    const int r_begin = first ? 0 : cg, r_inc = first ? 1 : nc_i;
                    if (!first || cg == 0) {
                    mbar_wait(bar_part + ring_part.slot, ring_part.parity);   // TMA data landed
                    const float* st = s_stage + (size_t)ring_part.slot * TC * RS;
                    const float* arow = s_a + (size_t)a_buf * TC * RS + s0p;
                    float* orow = s_occ + (size_t)o_buf * TC * ER + wc * OW;
                    unsigned* ocl = reinterpret_cast<unsigned*>(orow);
                    float* obl = orow + VO + lane;
                    struct RowData { float aB[P], aY[P], pb[P], py[P]; int off, ob, oy; };
                    auto load_rd = [&](int r, RowData& d) {
                        // the chunk was staged in frame order: the reversed sweep walks it backwards
                        const float* str = st + (size_t)(rev ? rows - 1 - r : r) * RS;
                        const float* stp = str + (hasX ? X * PW : 0);        // partner thread's P blanks
                        const int* sto = reinterpret_cast<const int*>(str + 2 * NP) + (hasX ? X : 0);
                        {
                            float qb[P], qy[P];
                            load_row<P>(stp + NP, qy, HS);     // first: the shuffles below wait for these
                            d.ob = *sto;
                            load_row<P>(stp, qb, HS);
                            load_row<P>(arow + r * RS, d.aB, HS);
                            load_row<P>(arow + r * RS + NP, d.aY, HS);
                            d.off = *reinterpret_cast<const int*>(arow + r * RS + offd);
    #pragma unroll
                            for (int q = 0; q < P; ++q) d.pb[q] = qb[P - 1 - q];
    #pragma unroll
                            for (int q = 0; q + 1 < P; ++q) d.py[q] = qy[P - 2 - q];
                            // my last label pairs with the LAST label of partner thread X-1 = the thread
                            // lane+1 of my warp talks to; lane 31 reads it itself
                            float yl = __shfl_down_sync(0xffffffffu, qy[P - 1], 1);
                            int ol = __shfl_down_sync(0xffffffffu, d.ob, 1);
                            if (lane == 31 && hasX1) { yl = stp[NP + (P == 8 ? HS : 0) - 1]; ol = sto[-1]; }
                            d.py[P - 1] = yl;
                            d.oy = ol;
                        }
                    };
                    // returns true when `gq` (label occupancies, Q1.31) still has to be added to the class slots
                    auto combine_rd = [&](int r, const RowData& d, unsigned (&gq)[P], float& bsum_out) -> bool {
                        const int u = tt0 + r - i0;
                        const float (&aB)[P] = d.aB; const float (&aY)[P] = d.aY;
                        const float (&pb)[P] = d.pb; const float (&py)[P] = d.py;
                        const int off = d.off, ob = d.ob, oy = d.oy;
                        const bool in_win = (unsigned)u < win_cons;
                        float bsum = 0.f;
                        bool pending = false;
                        if (first && r == 0) {
                            // first combined row: also yields the likelihood P = sum_s a * p~.
                            // Exponent/mantissa form: E0 = max exponent of any term, z = sum of the
                            // terms scaled by 2^-E0; log2 P = E0 + log2 z.
                            float tB[P], tY[P];
                            int eB[P], eY[P];
    #pragma unroll
                            for (int q = 0; q < P; ++q) {
                                tB[q] = 0.f; tY[q] = 0.f; eB[q] = kLinNone; eY[q] = kLinNone;
                                if (in_win && vB[q] && aB[q] > 0.f && pb[q] > 0.f) {
                                    const int ea = expo(aB[q]), ep = expo(pb[q]);
                                    tB[q] = (aB[q] * pow2c(-ea)) * (pb[q] * pow2c(-ep));
                                    eB[q] = ea + ep + off + ob;
                                }
                                const bool okY = q + 1 < P ? true : hasX1;
                                const int oq = q + 1 < P ? ob : oy;
                                if (in_win && vY[q] && okY && aY[q] > 0.f && py[q] > 0.f) {
                                    const int ea = expo(aY[q]), ep = expo(py[q]);
                                    tY[q] = (aY[q] * pow2c(-ea)) * (py[q] * pow2c(-ep));
                                    eY[q] = ea + ep + off + oq;
                                }
                            }
                            int em = kLinNone;
    #pragma unroll
                            for (int q = 0; q < P; ++q) em = max(em, max(eB[q], eY[q]));
                            em = warp_max_i(em);
                            if (R > 1) {
                                if (lane == 0) s_red[64 + wc] = em;
                                named_bar_sync(2, nbar);
                                for (int i = 0; i < R; ++i) em = max(em, s_red[64 + i]);
                                named_bar_sync(2, nbar);
                            }
                            const bool dead = em == kLinNone;      // no path survived (infeasible or underflow)
                            E0 = dead ? 0 : em;
                            float z = 0.f;
    #pragma unroll
                            for (int q = 0; q < P; ++q) {
                                tB[q] = eB[q] == kLinNone ? 0.f : tB[q] * pow2c(eB[q] - E0);
                                tY[q] = eY[q] == kLinNone ? 0.f : tY[q] * pow2c(eY[q] - E0);
                                z += tB[q] + tY[q];
                            }
                            z = warp_sum(z);
                            if (R > 1) {
                                if (lane == 0) s_red[128 + wc] = __float_as_int(z);
                                named_bar_sync(2, nbar);
                                z = 0.f;
                                for (int i = 0; i < R; ++i) z += __int_as_float(s_red[128 + i]);
                                named_bar_sync(2, nbar);
                            }
                            const bool bad = dead || !(z > 0.f) || !(z < 3.0e38f);
                            rz = bad ? 0.f : 1.0f / z;
                            if (tid == 0) {
                                s_red[100] = E0;
                                s_red[101] = __float_as_int(rz);
                                if (bad) { s_flag[0] = 1; s_flag[1] = 1; }
                                if (!rev) p.nll[b] = bad ? 0.f : (float)(-((double)E0 + (double)log2f(z)) * kLn2);
                            }
                            if (wgc) {
    #pragma unroll
                                for (int q = 0; q < P; ++q) {
                                    bsum += tB[q] * rz;
                                    atomicAdd(occ_slot(ocl + r * ER, lab[q]), __float2uint_rn(tY[q] * (rz * kQ31)));
                                }
                            }
                        } else if (in_win) {
                            // occupancy = a * p~ * 2^(off + o - E0) / z   (exact exponents); cells outside
                            // [0, S] are exact zeros on REC's side, partner vectors are finite
                            // (label cells directly in Q1.31 units)
                            // The exponent h = off + o - E0 goes onto the partner cell up to kLinHmax (so that
                            // p * 2^h cannot overflow); a larger h -- cells far below their thread's maximum on
                            // both sides, steep lattices -- puts the rest onto the product (a factor of 1 otherwise).
                            const int hb = off + ob - E0, hy = off + oy - E0;
                            const int hbc = max(min(hb, kLinHmax), kLinHmin), hyc = max(min(hy, kLinHmax), kLinHmin);
                            const float sb = pow2c(hbc) * rz;
                            const float sq = sb * kQ31;
                            const float sy = hasX1 ? pow2c(hyc) * (rz * kQ31) : 0.f;
                            const float rb = pow2c(hb - hbc), ry = pow2c(hy - hyc);   // 1 unless h is outside [kLinHmin, kLinHmax]
    #pragma unroll
                            for (int q = 0; q < P; ++q) {
                                bsum += (aB[q] * (pb[q] * sb)) * rb;
                                float gy = 0.f;
                                if (q + 1 < P) gy = (aY[q] * (py[q] * sq)) * rb;
                                else if (hasX1) gy = (aY[q] * (py[q] * sy)) * ry;
                                gq[q] = __float2uint_rn(gy);
                            }
                            pending = true;
                        }
                        bsum_out = bsum;
                        return pending;
                    };
                    // Steady state with registers to spare (4-warp CTAs): TWO rows per pass, staged so that
                    // at most ~48 cell registers are live: label planes of both rows -> Q1.31 occupancies;
                    // blank planes of both rows -> blank sums; then all atomics.  Two independent dependency
                    // chains per warp instead of one.
                    auto combine2 = [&](int rA, int rB) {
                        const float* strA = st + (size_t)(rev ? rows - 1 - rA : rA) * RS;
                        const float* strB = st + (size_t)(rev ? rows - 1 - rB : rB) * RS;
                        const float* stpA = strA + (hasX ? X * PW : 0);
                        const float* stpB = strB + (hasX ? X * PW : 0);
                        const int* stoA = reinterpret_cast<const int*>(strA + 2 * NP) + (hasX ? X : 0);
                        const int* stoB = reinterpret_cast<const int*>(strB + 2 * NP) + (hasX ? X : 0);
                        const float* arA = arow + rA * RS;
                        const float* arB = arow + rB * RS;
                        float qyA[P], qyB[P], aYA[P], aYB[P];
                        load_row<P>(stpA + NP, qyA, HS);
                        load_row<P>(stpB + NP, qyB, HS);
                        const int obA = *stoA, obB = *stoB;
                        const int offA = *reinterpret_cast<const int*>(arA + offd);
                        const int offB = *reinterpret_cast<const int*>(arB + offd);
                        load_row<P>(arA + NP, aYA, HS);
                        load_row<P>(arB + NP, aYB, HS);
                        float ylA = __shfl_down_sync(0xffffffffu, qyA[P - 1], 1);
                        float ylB = __shfl_down_sync(0xffffffffu, qyB[P - 1], 1);
                        int olA = __shfl_down_sync(0xffffffffu, obA, 1);
                        int olB = __shfl_down_sync(0xffffffffu, obB, 1);
                        if (lane == 31 && hasX1) {
                            ylA = stpA[NP + (P == 8 ? HS : 0) - 1]; olA = stoA[-1];
                            ylB = stpB[NP + (P == 8 ? HS : 0) - 1]; olB = stoB[-1];
                        }
                        const bool winA = (unsigned)(tt0 + rA - i0) < win_cons;
                        const bool winB = (unsigned)(tt0 + rB - i0) < win_cons;
                        // occupancy = a * p~ * 2^(off + o - E0) / z (see combine_rd for the exponent split)
                        const int hbA = offA + obA - E0, hyA = offA + olA - E0;
                        const int hbB = offB + obB - E0, hyB = offB + olB - E0;
                        const int cbA = max(min(hbA, kLinHmax), kLinHmin), cyA = max(min(hyA, kLinHmax), kLinHmin);
                        const int cbB = max(min(hbB, kLinHmax), kLinHmin), cyB = max(min(hyB, kLinHmax), kLinHmin);
                        const float sbA = winA ? pow2c(cbA) * rz : 0.f, sbB = winB ? pow2c(cbB) * rz : 0.f;
                        const float syA = (winA && hasX1) ? pow2c(cyA) * (rz * kQ31) : 0.f;
                        const float syB = (winB && hasX1) ? pow2c(cyB) * (rz * kQ31) : 0.f;
                        const float rbA = pow2c(hbA - cbA), ryA = pow2c(hyA - cyA);
                        const float rbB = pow2c(hbB - cbB), ryB = pow2c(hyB - cyB);
                        const float sqA = sbA * kQ31, sqB = sbB * kQ31;
                        unsigned gA[P], gB[P];
    #pragma unroll
                        for (int q = 0; q + 1 < P; ++q) {
                            gA[q] = __float2uint_rn((aYA[q] * (qyA[P - 2 - q] * sqA)) * rbA);
                            gB[q] = __float2uint_rn((aYB[q] * (qyB[P - 2 - q] * sqB)) * rbB);
                        }
                        gA[P - 1] = __float2uint_rn((aYA[P - 1] * (ylA * syA)) * ryA);
                        gB[P - 1] = __float2uint_rn((aYB[P - 1] * (ylB * syB)) * ryB);
                        float bsA = 0.f, bsB = 0.f;
                        {
                            float qbA[P], qbB[P], aBA[P], aBB[P];
                            load_row<P>(stpA, qbA, HS);
                            load_row<P>(stpB, qbB, HS);
                            load_row<P>(arA, aBA, HS);
                            load_row<P>(arB, aBB, HS);
    #pragma unroll
                            for (int q = 0; q < P; ++q) {
                                bsA += (aBA[q] * (qbA[P - 1 - q] * sbA)) * rbA;
                                bsB += (aBB[q] * (qbB[P - 1 - q] * sbB)) * rbB;
                            }
                        }
                        if (winA) {
    #pragma unroll
                            for (int q = 0; q < P; ++q) atomicAdd(occ_slot(ocl + rA * ER, lab[q]), gA[q]);
                        }
                        if (winB) {
    #pragma unroll
                            for (int q = 0; q < P; ++q) atomicAdd(occ_slot(ocl + rB * ER, lab[q]), gB[q]);
                        }
                        // (a thread outside its window may have read lattice cells nobody wrote: whatever they held
                        // -- NaN bit patterns of a recycled allocation included -- must not reach the blank sum)
                        obl[rA * ER] = winA ? bsA : 0.f;
                        obl[rB * ER] = winB ? bsB : 0.f;
                    };
                    int r = r_begin;
                    // (two rows in flight need ~128 registers per thread: the 4-warp CTAs, and every instantiation whose
                    // launch bounds leave them -- MAXT * MINB <= 512)
                    // (WIDE: chunks of 2 frames, one row per combine warp)
                    if (!WIDE && !first && wgc && (NT <= 128 || MAXT * MINB <= 512))
                        for (; r + r_inc < rows; r += 2 * r_inc) combine2(r, r + r_inc);
                    for (; r < rows; r += r_inc) {
                        RowData d0;
                        unsigned gq[P];
                        float bsum;
                        load_rd(r, d0);
                        if (combine_rd(r, d0, gq, bsum)) {
    #pragma unroll
                            for (int q = 0; q < P; ++q) atomicAdd(occ_slot(ocl + r * ER, lab[q]), gq[q]);
                        }
                        if (wgc) obl[r * ER] = bsum;
                    }
                    }
                    ring_part.advance();
                    a_buf ^= 1;
                    o_buf ^= 1;
                }
                LPROF_END(it >= n1 + 2);
                __syncthreads();
                if (it == n1) {
                    // Phase break: my REC warps have stored every row the partner will consume,
                    // and (after the cluster barrier) vice versa.
                    cluster_sync_all();
                    if (iss_part) {
                        fence_proxy_async();
                        for (int kk = n1; kk < n1 + kLinPDist; ++kk) {
                            if (kk < nch && (want_grad || kk == n1)) issue_partner(kk, iss_p.slot);
                            iss_p.advance();
                        }
                    }
                    __syncthreads();
                }
        };
        int it = 0;
        if constexpr (RCL) {
            // Steady state of the headline shape class in a loop of its own: chunks k = it - 2 in (n1, nch - 1), i.e.
            // neither the first combined chunk nor the (possibly short) last one; full chunks of 4 rows, two
            // combine groups (this warp takes rows cg and cg + 2).  Every shared-memory address is an immediate
            // offset from a running 32-bit base; what only the general path needs is not live in here.
            static_assert(!RCL || kLinPDist >= 2, "the steady-state loop requests the partner rows kLinPDist chunks ahead");
            static_assert(!RISS || (FIX && kLinPDist == 2), "the recursion warp's requests assume a lead of two chunks");
            const int it_fast_end = wgc ? nch_i + 1 : 0;
            if (wgc) {
                // First half: the recursion warp publishes the pre-emission rows of chunk it - 1 in the
                // shared-memory ring (full chunks 0 .. n1 - 2); this warp copies rows cg and cg + 2 of chunk
                // it - 2 to the lattice in HBM, inside the same band of cells the recursion warp would store.
                const unsigned win_copy = (i0 - P <= S) ? (unsigned)(C + 3 * P) : 0u;
                const unsigned cp0 = sbase + lay.a + (unsigned)lane * 16u;
                const unsigned exo = 2048u + (unsigned)lane * 4u - (unsigned)lane * 16u;
                auto copy_chunk = [&](int k) {
                    const unsigned src = cp0 + (unsigned)(k & 1) * (4u * 544u * 4u);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int rr = cg + 2 * h, tt = 4 * k + rr;
                        if ((unsigned)(tt - i0 + P) < win_copy) {
                            const unsigned a = src + (unsigned)rr * (544u * 4u);
                            const float4 v0 = lds128(a), v1 = lds128(a + 512u), v2 = lds128(a + 1024u), v3 = lds128(a + 1536u);
                            const int e = lds32i(a + exo);
                            float* dst = lat_b + (ptrdiff_t)(tbase + tsign * tt) * 544 + lane * 4;
                            *reinterpret_cast<float4*>(dst) = v0;
                            *reinterpret_cast<float4*>(dst + 128) = v1;
                            *reinterpret_cast<float4*>(dst + 256) = v2;
                            *reinterpret_cast<float4*>(dst + 384) = v3;
                            *reinterpret_cast<int*>(dst + 512 + lane - lane * 4) = e;
                        }
                    }
                };
#if CTC_LIN_P1RING
                for (; it < n1_i; ++it) {
                    if (it >= 2) copy_chunk(it - 2);
                    __syncthreads();
                }
                if (it == n1_i && n1_i >= 2) copy_chunk(n1_i - 2);
#endif
            }
            for (; it < n_it && it < n1_i + 3; ++it) comb_iter(it);
            if (it < it_fast_end) {
                if (cg > 0) { E0 = s_red[100]; rz = __int_as_float(s_red[101]); }
                unsigned uA = (unsigned)(n_store + (it - 2 - n1_i) * 4 + cg - i0);    // sweep step of row A minus my first pair
                const unsigned bar0 = sbase + lay.bars + 8u * (unsigned)lay.NL;       // bar_part[0]
                for (; it < it_fast_end; ++it) {
                    LPROF_BEGIN();
                    if (it + kLinPDist - 2 < nch_i) {     // partner rows of chunk it + kLinPDist - 2 (consumed in iteration it + kLinPDist)
                        if (!RISS && iss_part && lane == 0) {   // (RISS: the recursion warp requests them)
                            const int tt0i = n_store + (it + kLinPDist - 2 - n1_i) * 4, rowsi = min(4, Tb - tt0i);
                            const int t_lo = rev ? tbase - (tt0i + rowsi - 1) : tt0i;
                            const unsigned bytes = (unsigned)rowsi * (544u * 4u), bar = bar0 + 8u * (unsigned)iss_p.slot;
                            mbar_expect_tx_a(bar, bytes);
                            bulk_g2s_a(sbase + lay.stage + (unsigned)iss_p.slot * (4u * 544u * 4u),
                                       lat_b + (ptrdiff_t)t_lo * 544, bytes, bar);
                        }
                        iss_p.advance();
                    }
                    mbar_wait_a(bar0 + 8u * (unsigned)ring_part.slot, ring_part.parity);   // TMA data landed
                    {
                        constexpr unsigned RSB = 544u * 4u, PLB = 256u * 4u, HSB = 128u * 4u;
                        const unsigned ERB = FIX ? 112u * 4u : (unsigned)OW * 4u;
                        const unsigned stb = sbase + lay.stage + (unsigned)ring_part.slot * (4u * RSB);
                        const unsigned stA = stb + fx_stA, stB = stb + fx_stB;
                        const unsigned arA = sbase + lay.a + (unsigned)a_buf * (4u * RSB) + fx_ar, arB = arA + 2u * RSB;
                        const unsigned ocA = sbase + lay.occ + (unsigned)o_buf * (4u * ERB) + (unsigned)cg * ERB, ocB = ocA + 2u * ERB;
                        float4 t0, t1;
                        float qyA[P], qyB[P], aYA[P], aYB[P];
#define CTC_LD8(dst, addr) t0 = lds128(addr); t1 = lds128((addr) + HSB); \
    dst[0] = t0.x; dst[1] = t0.y; dst[2] = t0.z; dst[3] = t0.w; dst[4] = t1.x; dst[5] = t1.y; dst[6] = t1.z; dst[7] = t1.w
                        CTC_LD8(qyA, stA + PLB);
                        CTC_LD8(qyB, stB + PLB);
                        const int obA = lds32i(stA + fx_e), obB = lds32i(stB + fx_e);
                        const int offA = lds32i(arA + fx_o), offB = lds32i(arB + fx_o);
                        CTC_LD8(aYA, arA + PLB);
                        CTC_LD8(aYB, arB + PLB);
                        float ylA = __shfl_down_sync(0xffffffffu, qyA[P - 1], 1);
                        float ylB = __shfl_down_sync(0xffffffffu, qyB[P - 1], 1);
                        int olA = __shfl_down_sync(0xffffffffu, obA, 1);
                        int olB = __shfl_down_sync(0xffffffffu, obB, 1);
                        if (lane == 31 && hasX1) {
                            ylA = lds32(stA + PLB + HSB - 4u); olA = lds32i(stA + fx_e - 4u);
                            ylB = lds32(stB + PLB + HSB - 4u); olB = lds32i(stB + fx_e - 4u);
                        }
                        const bool winA = uA < win_cons;
                        const bool winB = uA + 2u < win_cons;
                        const int hbA = offA + obA - E0, hyA = offA + olA - E0;
                        const int hbB = offB + obB - E0, hyB = offB + olB - E0;
                        const int cbA = max(min(hbA, kLinHmax), kLinHmin), cyA = max(min(hyA, kLinHmax), kLinHmin);
                        const int cbB = max(min(hbB, kLinHmax), kLinHmin), cyB = max(min(hyB, kLinHmax), kLinHmin);
                        const float sbA = winA ? pow2c(cbA) * rz : 0.f, sbB = winB ? pow2c(cbB) * rz : 0.f;
                        const float syA = (winA && hasX1) ? pow2c(cyA) * (rz * kQ31) : 0.f;
                        const float syB = (winB && hasX1) ? pow2c(cyB) * (rz * kQ31) : 0.f;
                        const float rbA = pow2c(hbA - cbA), ryA = pow2c(hyA - cyA);
                        const float rbB = pow2c(hbB - cbB), ryB = pow2c(hyB - cyB);
                        const float sqA = sbA * kQ31, sqB = sbB * kQ31;
                        unsigned gA[P], gB[P];
#pragma unroll
                        for (int q = 0; q + 1 < P; ++q) {
                            gA[q] = __float2uint_rn((aYA[q] * (qyA[P - 2 - q] * sqA)) * rbA);
                            gB[q] = __float2uint_rn((aYB[q] * (qyB[P - 2 - q] * sqB)) * rbB);
                        }
                        gA[P - 1] = __float2uint_rn((aYA[P - 1] * (ylA * syA)) * ryA);
                        gB[P - 1] = __float2uint_rn((aYB[P - 1] * (ylB * syB)) * ryB);
                        float bsA = 0.f, bsB = 0.f;
                        {
                            float qbA[P], qbB[P], aBA[P], aBB[P];
                            CTC_LD8(qbA, stA);
                            CTC_LD8(qbB, stB);
                            CTC_LD8(aBA, arA);
                            CTC_LD8(aBB, arB);
#pragma unroll
                            for (int q = 0; q < P; ++q) {
                                bsA += (aBA[q] * (qbA[P - 1 - q] * sbA)) * rbA;
                                bsB += (aBB[q] * (qbB[P - 1 - q] * sbB)) * rbB;
                            }
                        }
#undef CTC_LD8
                        if (winA) {
#pragma unroll
                            for (int q = 0; q < P; ++q) reds_add_u32(ocA + (unsigned)lab[q], gA[q]);
                        }
                        if (winB) {
#pragma unroll
                            for (int q = 0; q < P; ++q) reds_add_u32(ocB + (unsigned)lab[q], gB[q]);
                        }
                        sts32(ocA + fx_bl, winA ? bsA : 0.f);
                        sts32(ocB + fx_bl, winB ? bsB : 0.f);
                    }
                    uA += 4u;
                    ring_part.advance();
                    a_buf ^= 1;
                    o_buf ^= 1;
                    LPROF_END(true);
                    __syncthreads();
                }
            }
        }
        for (; it < n_it; ++it) comb_iter(it);
    } else {
        // =============================================================================
        // SOFT / GRAD: logits staging + fused softmax; gradient rows
        // =============================================================================
        // A helper owns F = TC / n frames of every chunk and works on all of them AT ONCE, a group
        // of G = 32 / F lanes per frame, so that one pass of short shuffle trees finishes the chunk.
        // (all of it by shifts: the compiler does not hoist integer divisions out of the chunk loop)
        const int lgTC = TC >= 3 ? 2 : (TC == 2 ? 1 : 0);   // TC = 3: four frame slots, the last one idle
        const int lgA = max(lgTC - (nA >= 4 ? 2 : (nA >= 2 ? 1 : 0)), 0), lgB = max(lgTC - (nB >= 4 ? 2 : (nB >= 2 ? 1 : 0)), 0);
        const int FA = 1 << lgA, FB = 1 << lgB;
        const int GA = 32 >> lgA, glA = lane & (GA - 1), fA = ha * FA + (lane >> (5 - lgA));
        const int GB = 32 >> lgB, glB = lane & (GB - 1), fB = hb * FB + (lane >> (5 - lgB));
        const unsigned gmaskA = (GA == 32 ? 0xffffffffu : ((1u << GA) - 1u)) << (lane & ~(GA - 1));
        // rows that are not 16-byte aligned in HBM: 16-byte copies + register-resident passes when a lane holds at
        // most 16 classes of its frame (V <= 256 with four helpers, V <= 128 with one)
        const bool shifted = MID && !al;
        // floats between the 16-byte line and the logits row of sweep step tt: (mis0 + tt * mis_step) & 3
        const int mis0 = (int)(((reinterpret_cast<uintptr_t>(acts_b) >> 2) + (uintptr_t)tbase * (uintptr_t)frame_stride) & 3);
        const int mis_step = (int)((uintptr_t)((ptrdiff_t)tsign * (ptrdiff_t)frame_stride) & 3);
        auto row_mis = [&](int tt) { return (mis0 + tt * mis_step) & 3; };
        auto group_sum = [&](float x, int G) {
            for (int o = G >> 1; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            return x;
        };
        auto group_max = [&](float x, int G) {
            for (int o = G >> 1; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
            return x;
        };

        // ---- logits rows of a chunk: cp.async, 16 B per lane and copy; a lane's (row, column) of
        //      its first two copies never change, so they are computed once
        const ptrdiff_t a_inc = (ptrdiff_t)tsign * (ptrdiff_t)frame_stride;
        // MID: softmax warp ha copies its own FA frames of a chunk (rows [r_lo, r_lo + r_cnt)) and waits for them
        // with cp.async groups; otherwise one warp copies the chunk (the mbarrier arrive that several consumers
        // need costs the issuing warp about a microsecond on B200)
        const bool own_rows = MID && nA > 1 && !wide_rows;
        const int r_lo = own_rows ? ha * FA : 0, r_cnt = own_rows ? FA : TC;
        const int n4 = r_cnt * V4;
        const bool cp_groups = (nA == 1 || own_rows) && !wide_rows;   // a softmax warp waits for its own cp.async groups
        int cp_dst[2], cp_row[2];
        bool pf_lane[2];
        ptrdiff_t cp_src[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int idx = lane + 32 * j, r = idx / max(V4, 1), c = idx - r * V4;
            cp_row[j] = idx < n4 ? r_lo + r : 0x7fffffff;
            pf_lane[j] = c == 0 || c == 8;   // two of a V = 48 row's lanes touch both of its 128-byte lines
            cp_dst[j] = (r_lo + r) * Vs + 4 * c;
            cp_src[j] = (r_lo + r) * a_inc + 4 * c;
        }
        auto issue_logits = [&](int ka, int slot_a) {
            int tt0, rows;
            chunk_at(ka, tt0, rows);
            float* dst = s_y + (size_t)slot_a * TC * Vs;
            const float* src = acts_b + (ptrdiff_t)(tbase + tsign * tt0) * (ptrdiff_t)frame_stride;
            if (wide_rows) {
                if (lane == 0) issue_logits_tma(ka, slot_a);
                return;
            }
            if (MID && !al && shifted) {
                // rows that start 4, 8 or 12 bytes into a 16-byte line (V % 4 != 0): whole 16-byte segments of the
                // lines the row touches; class c lands at ring position mis + c (the bytes in front of the row
                // belong to the previous row of the tensor, the tail of the last segment is zero-filled).  At most
                // 65 segments per row (V <= 256): three predicated copies per lane and row, straight-line.
#pragma unroll
                // (shared memory by explicit 32-bit addresses off the pinned base: a generic pointer costs an
                // S2R SR_CgaCtaId + LEA per use in a kernel with cluster dimensions)
                const unsigned dst_a = sbase + (unsigned)lay.y + (unsigned)(slot_a * TC * Vs + 4 * lane) * 4u;
#pragma unroll
                for (int rr = 0; rr < (MID8 ? 1 : 2); ++rr) {      // MID: this warp's two frames of the chunk (MID8: its frame)
                    const int r = r_lo + rr;
                    if (r < rows) {
                        const int mis = (mis0 + (tt0 + r) * mis_step) & 3;
                        const float* s16 = src + r * a_inc - mis + 4 * lane;      // 16-byte aligned
                        const unsigned d16 = dst_a + (unsigned)(r * Vs) * 4u;
                        const int left = (mis + V) * 4 - 16 * lane;             // bytes of the row from my first segment on
#pragma unroll
                        for (int q = 0; q < 3; ++q)
                            if (left - 512 * q > 0) cp_async16_zfill_a(d16 + 512u * q, s16 + 128 * q, min(16, left - 512 * q));
                    }
                }
            } else if (!al) {
                for (int r = r_lo; r < min(rows, r_lo + r_cnt); ++r)
                    for (int c = lane; c < V; c += 32) cp_async4(dst + r * Vs + c, src + r * a_inc + c);
            } else {
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (cp_row[j] < rows) cp_async16(dst + cp_dst[j], src + cp_src[j]);
            }
            if (al && n4 > 64) {
                const int r_end = min(rows, r_lo + r_cnt);
                int r = r_lo + 64 / V4, c = 64 - (64 / V4) * V4 + lane;
                while (c >= V4) { c -= V4; ++r; }
                while (r < r_end) {
                    cp_async16(dst + r * Vs + 4 * c, src + r * a_inc + 4 * c);
                    c += 32;
                    while (c >= V4) { c -= V4; ++r; }
                }
            }
            // one softmax warp: it waits for its own copies with cp.async groups (the mbarrier arrive
            // costs the issuing warp about a microsecond on B200); several: completion on the mbarrier
            if (cp_groups) cp_async_commit();
            else cp_async_arrive(bar_acts + slot_a);
        };

        // V = 48 with four frames per pass: straight-line code, the row stays in registers
        auto softmax_fast = [&](auto CL, float* base, int rows, const float2 (&lg)[3]) {
                constexpr bool CLAMPED = decltype(CL)::value;   // compile-time copy of clp.on (uniform branch at the call)
                const Clamp clq{CLAMPED, clp.lo, clp.hi};
                (void)clq;
            const bool act = fA < rows;
            float* row = base + min(fA, rows - 1) * Vs;
            float2* row2 = reinterpret_cast<float2*>(row);
            float2 x[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) x[j] = act ? lg[j] : make_float2(0.f, 0.f);
            unsigned mk = 0u;                  // fused Hardtanh: bit 2j / 2j+1 = gradient blocked
            if (clq.on) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    mk |= (clq.cmask(x[j].x) ? 1u : 0u) << (2 * j) | (clq.cmask(x[j].y) ? 1u : 0u) << (2 * j + 1);
                    x[j].x = clq.cin(x[j].x);
                    x[j].y = clq.cin(x[j].y);
                }
            }
            float m = fmaxf(fmaxf(fmaxf(x[0].x, x[0].y), fmaxf(x[1].x, x[1].y)), fmaxf(x[2].x, x[2].y));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            const float mb = m * kLog2e;
            float z = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                x[j].x = ex2f(fmaf(x[j].x, kLog2e, -mb));
                x[j].y = ex2f(fmaf(x[j].y, kLog2e, -mb));
                z += x[j].x + x[j].y;
            }
            // (redux.sync was measured SLOWER than three shuffle levels here, for the sum and the max)
            z += __shfl_xor_sync(0xffffffffu, z, 4);
            z += __shfl_xor_sync(0xffffffffu, z, 2);
            z += __shfl_xor_sync(0xffffffffu, z, 1);
            float rs;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(z));
            rs = rs * (2.0f - z * rs);          // one Newton step: full fp32 accuracy
            if (act) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float a = x[j].x * rs, c = x[j].y * rs;
                    row2[glA + 8 * j] = make_float2((mk >> (2 * j)) & 1u ? -a : a, (mk >> (2 * j + 1)) & 1u ? -c : c);
                }
                if (glA == 0) row[V] = 0.f;     // what padding pairs gather
            }
        };

        // ---- fused softmax, in place, of my F frames of a chunk (a group of G lanes per frame) ----
        // The maximum of a row comes from ONE redux.sync per group (no shuffle tree); the sum needs
        // log2 G shuffle levels.
        auto softmax_chunk = [&](auto CL, float* base, int rows, int tt0) {
                constexpr bool CLAMPED = decltype(CL)::value;   // compile-time copy of clp.on (uniform branch at the call)
                const Clamp clq{CLAMPED, clp.lo, clp.hi};
                (void)clq;
            const int G = GA, gl = glA, f = fA;
            const unsigned gmask = gmaskA;
            const bool act = f < rows;
            float* row = base + min(f, rows - 1) * Vs;
            float2* row2 = reinterpret_cast<float2*>(row);
            if constexpr (MID) {
                // MID: chunks of 4 frames, two softmax warps, 16 lanes per frame: lane gl holds classes gl, gl + 16, ...
                // (at most NV = 12 of them up to V = 192, 16 up to 256) in registers, loaded by immediate offsets.  A row
                // that is not 16-byte aligned in HBM sits `mis` floats into its ring row and is moved to the front here.
                auto body = [&](auto NVC) {
                    constexpr int NV = decltype(NVC)::value;
                    // (explicit 32-bit shared-memory addresses, as in the FIX loops)
                    const unsigned ra = sbase + (unsigned)lay.y + (unsigned)((int)(base - s_y) + min(f, rows - 1) * Vs + gl) * 4u;
                    const unsigned rin = ra + (unsigned)(shifted ? row_mis(tt0 + min(f, rows - 1)) : 0) * 4u;
                    const int nv = (V - gl + MG - 1) / MG;      // classes this lane holds
                    float x[NV];
                    float m = -CUDART_INF_F;
                    unsigned mk = 0u;
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        // (branch-free: every lane loads NV values -- beyond V whatever the ring holds there -- and selects)
                        const float raw = lds32(rin + (unsigned)(MG * 4 * j));
                        const bool on = j < nv;
                        mk |= ((on && clq.cmask(raw)) ? 1u : 0u) << j;
                        x[j] = on ? clq.cin(raw) : -CUDART_INF_F;
                        m = fmaxf(m, x[j]);
                    }
                    if constexpr (MG == 32 && CTC_LIN_REDUX) {      // a warp per frame: one REDUX (five shuffle levels otherwise)
                        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(m));
                    } else {
#pragma unroll
                        for (int o = MG / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                    }
                    const float mb = m * kLog2e;
                    float z = 0.f;
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        x[j] = ex2f(fmaf(x[j], kLog2e, -mb));      // (-inf for the classes beyond V: 0)
                        z += x[j];
                    }
#pragma unroll
                    for (int o = MG / 2; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
                    float rs;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(z));
                    rs = rs * (2.0f - z * rs);          // one Newton step: full fp32 accuracy (1 <= z <= V)
                    __syncwarp();      // every lane has read its raw values before the row is overwritten in place
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        const float y = x[j] * rs;
                        sts32_if(ra + (unsigned)(MG * 4 * j), (mk >> j) & 1u ? -y : y, act && j < nv);
                    }
                    sts32_if(ra + (unsigned)V * 4u, 0.f, act && gl < Vs - V);   // slot V (and the padding behind it): what padding pairs gather
                };
                // (two buckets only: a third one for V <= 64 cost the four-helper launches 5 ... 9 % -- instruction fetch)
                if (V <= 192) body(std::integral_constant<int, 192 / MG>{});
                else body(std::integral_constant<int, 256 / MG>{});
                return;
            }
            if (!WIDE && !al) {      // rows that are not 16-byte aligned in HBM (V % 4 != 0): scalar passes
                float m = -CUDART_INF_F, z = 0.f;
                for (int c = gl; c < V; c += G) m = fmaxf(m, clq.cin(row[c]));
                m = group_max(m, G);
                for (int c = gl; c < V; c += G) z += ex2f((clq.cin(row[c]) - m) * kLog2e);
                const float rs = 1.0f / group_sum(z, G);
                if (act) {
                    for (int c = gl; c < V; c += G) {
                        const float raw = row[c], y = ex2f((clq.cin(raw) - m) * kLog2e) * rs;
                        row[c] = clq.cmask(raw) ? -y : y;
                    }
                    for (int c = V + gl; c < Vs; c += G) row[c] = 0.f;   // slot V: what padding pairs gather
                }
                return;
            }
            if (!WIDE && G == 8 && V2 == 24) {      // V = 48, four frames per pass: straight-line code, no guards
                float2 lg[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) lg[j] = row2[gl + 8 * j];
                softmax_fast(CL, base, rows, lg);
                return;
            }
            if (!WIDE && V2 <= 4 * G) {      // at most 4 float2 per lane: the row stays in registers
                float2 x[4];
                float m = -CUDART_INF_F;
                unsigned mk = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = gl + j * G;
                    x[j] = c < V2 ? row2[c] : make_float2(-CUDART_INF_F, -CUDART_INF_F);
                    if (clq.on && c < V2) {
                        mk |= (clq.cmask(x[j].x) ? 1u : 0u) << (2 * j) | (clq.cmask(x[j].y) ? 1u : 0u) << (2 * j + 1);
                        x[j].x = clq.cin(x[j].x);
                        x[j].y = clq.cin(x[j].y);
                    }
                    m = fmaxf(m, fmaxf(x[j].x, x[j].y));
                }
                asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(m) : "f"(m), "r"(gmask));
                const float mb = m * kLog2e;
                float z = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    x[j].x = ex2f(fmaf(x[j].x, kLog2e, -mb));
                    x[j].y = ex2f(fmaf(x[j].y, kLog2e, -mb));
                    z += x[j].x + x[j].y;
                }
                const float rs = 1.0f / group_sum(z, G);
                if (act) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = gl + j * G;
                        if (c < V2) {
                            const float a = x[j].x * rs, d = x[j].y * rs;
                            row2[c] = make_float2((mk >> (2 * j)) & 1u ? -a : a, (mk >> (2 * j + 1)) & 1u ? -d : d);
                        }
                    }
                }
            } else if (WIDE && CTC_LIN_WIDE_FAST && V4 <= 8 * 32) {
                // WIDE, a warp per frame, at most 8 x 128 bit per lane (C4): as the branch below, but branch-free and by
                // explicit 32-bit shared addresses (every lane loads 8 x 128 bit -- beyond V whatever the ring holds there --
                // and selects; predicated stores), row maximum by ONE redux.sync.max.f32, rcp + Newton step
                const unsigned ra = sbase + (unsigned)lay.y + (unsigned)((int)(base - s_y) + min(f, rows - 1) * Vs) * 4u + (unsigned)gl * 16u;
                float4 x[8];
                float m = -CUDART_INF_F;
                unsigned mk = 0u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool on = gl + j * 32 < V4;
                    const float4 raw = lds128(ra + (unsigned)(j * 512));
                    if (clq.on)
                        mk |= ((clq.cmask(raw.x) ? 1u : 0u) | (clq.cmask(raw.y) ? 2u : 0u) |
                               (clq.cmask(raw.z) ? 4u : 0u) | (clq.cmask(raw.w) ? 8u : 0u)) << (4 * j);
                    x[j].x = on ? clq.cin(raw.x) : -CUDART_INF_F; x[j].y = on ? clq.cin(raw.y) : -CUDART_INF_F;
                    x[j].z = on ? clq.cin(raw.z) : -CUDART_INF_F; x[j].w = on ? clq.cin(raw.w) : -CUDART_INF_F;
                    m = fmaxf(m, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
                }
                asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(m));
                const float mb = m * kLog2e;
                float z0 = 0.f, z1 = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    x[j].x = ex2f(fmaf(x[j].x, kLog2e, -mb));
                    x[j].y = ex2f(fmaf(x[j].y, kLog2e, -mb));
                    x[j].z = ex2f(fmaf(x[j].z, kLog2e, -mb));
                    x[j].w = ex2f(fmaf(x[j].w, kLog2e, -mb));
                    z0 += x[j].x + x[j].y;
                    z1 += x[j].z + x[j].w;
                }
                const float z = group_sum(z0 + z1, 32);
                float rs;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(z));
                rs = rs * (2.0f - z * rs);          // one Newton step: full fp32 accuracy (1 <= z <= V)
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned q = mk >> (4 * j);
                    const float a = x[j].x * rs, d = x[j].y * rs, e = x[j].z * rs, h = x[j].w * rs;
                    sts128_if(ra + (unsigned)(j * 512), q & 1u ? -a : a, q & 2u ? -d : d, q & 4u ? -e : e, q & 8u ? -h : h,
                              act && gl + j * 32 < V4);
                }
            } else if (YS == 0 && V4 <= 8 * G) {
                // wide vocabulary (V = 1024 with a warp per frame): at most 8 float4 per lane, the row
                // stays in registers and every load is issued before the first use (the looped path
                // below pays the shared-memory latency once per element and pass)
                float4* row4 = reinterpret_cast<float4*>(row);
                float4 x[8];
                float m = -CUDART_INF_F;
                unsigned mk = 0u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = gl + j * G;
                    x[j] = c < V4 ? row4[c] : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
                    if (clq.on && c < V4) {
                        mk |= ((clq.cmask(x[j].x) ? 1u : 0u) | (clq.cmask(x[j].y) ? 2u : 0u) |
                               (clq.cmask(x[j].z) ? 4u : 0u) | (clq.cmask(x[j].w) ? 8u : 0u)) << (4 * j);
                        x[j].x = clq.cin(x[j].x); x[j].y = clq.cin(x[j].y);
                        x[j].z = clq.cin(x[j].z); x[j].w = clq.cin(x[j].w);
                    }
                    m = fmaxf(m, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
                }
                m = group_max(m, G);
                const float mb = m * kLog2e;
                float z0 = 0.f, z1 = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    x[j].x = ex2f(fmaf(x[j].x, kLog2e, -mb));
                    x[j].y = ex2f(fmaf(x[j].y, kLog2e, -mb));
                    x[j].z = ex2f(fmaf(x[j].z, kLog2e, -mb));
                    x[j].w = ex2f(fmaf(x[j].w, kLog2e, -mb));
                    z0 += x[j].x + x[j].y;
                    z1 += x[j].z + x[j].w;
                }
                const float rs = 1.0f / group_sum(z0 + z1, G);
                if (act) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = gl + j * G;
                        if (c < V4) {
                            const unsigned q = mk >> (4 * j);
                            const float a = x[j].x * rs, d = x[j].y * rs, e = x[j].z * rs, h = x[j].w * rs;
                            row4[c] = make_float4(q & 1u ? -a : a, q & 2u ? -d : d, q & 4u ? -e : e, q & 8u ? -h : h);
                        }
                    }
                }
            } else {
                float m = -CUDART_INF_F, z = 0.f;
                for (int c = gl; c < V2; c += G) {
                    const float2 x = row2[c];
                    m = fmaxf(m, fmaxf(clq.cin(x.x), clq.cin(x.y)));
                }
                asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(m) : "f"(m), "r"(gmask));
                for (int c = gl; c < V2; c += G) {
                    const float2 x = row2[c];
                    z += ex2f((clq.cin(x.x) - m) * kLog2e) + ex2f((clq.cin(x.y) - m) * kLog2e);
                }
                const float rs = 1.0f / group_sum(z, G);
                if (act)
                    for (int c = gl; c < V2; c += G) {
                        const float2 x = row2[c];
                        const float a = ex2f((clq.cin(x.x) - m) * kLog2e) * rs, d = ex2f((clq.cin(x.y) - m) * kLog2e) * rs;
                        row2[c] = make_float2(clq.cmask(x.x) ? -a : a, clq.cmask(x.y) ? -d : d);
                    }
            }
            if (act && gl == 0) row[V] = 0.f;  // what padding pairs gather
        };

        // ---- gradient rows of my F frames of a chunk (a group of G lanes per frame) ---------------
        //   occupancy row = per recursion warp [class sums, Q1.31: VO][blank partial sums, fp32: 32].
        //   The row is read, cleared for its next use, and turned into gscale * (softmax - occupancy).
        //   The sum of ALL occupancies of a frame must be 1: the posterior-mass check.
        //   All loads are issued up front and the two reductions (blank sum, total) share their
        //   shuffle levels, so that the pass is one short dependent chain.
        auto grad_chunk = [&](auto CL, float* obase, const float* ybase, int tt0, int rows) {
                constexpr bool CLAMPED = decltype(CL)::value;   // compile-time copy of clp.on (uniform branch at the call)
                const Clamp clq{CLAMPED, clp.lo, clp.hi};
                (void)clq;
            const int G = GB, gl = glB, f = fB;
            const bool act = f < rows;
            const int fr = min(f, rows - 1);
            float* orow = obase + fr * ER;
            const float2* y2 = reinterpret_cast<const float2*>(ybase + fr * Vs);
            float2* g2 = reinterpret_cast<float2*>(grad_b + (size_t)(tbase + tsign * (tt0 + fr)) * frame_stride);
            if (!WIDE && RC == 1 && G == 8 && V2 == 24) {   // V = 48, four frames per pass: straight-line code
                const float4 bp = *reinterpret_cast<const float4*>(orow + VO + 4 * gl);
                uint2 x[3];
                float2 y[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    uint2* p2 = reinterpret_cast<uint2*>(orow) + gl + 8 * j;
                    x[j] = *p2;
                    y[j] = y2[gl + 8 * j];
                    if (act) *p2 = make_uint2(0u, 0u);
                }
                float bs = (bp.x + bp.y) + (bp.z + bp.w), tot = 0.f;
                float2 o[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    o[j].x = __uint2float_rn(x[j].x) * (1.0f / kQ31);
                    o[j].y = __uint2float_rn(x[j].y) * (1.0f / kQ31);
                    tot += o[j].x + o[j].y;
                }
#pragma unroll
                for (int sft = 4; sft > 0; sft >>= 1) {
                    bs += __shfl_xor_sync(0xffffffffu, bs, sft);
                    tot += __shfl_xor_sync(0xffffffffu, tot, sft);
                }
                // the blank class sits in column (blank >> 1) of lane (blank >> 1) & 7, slot (blank >> 1) >> 3
                const int cb = blank >> 1;
                const float addx = (blank & 1) ? 0.f : bs, addy = (blank & 1) ? bs : 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const bool mine = cb == gl + 8 * j;
                    const float ox = o[j].x + (mine ? addx : 0.f), oy = o[j].y + (mine ? addy : 0.f);
                    // (a set sign bit of y: the fused Hardtanh blocks this entry's gradient)
                    if (act) g2[gl + 8 * j] = make_float2((CLAMPED && __float_as_int(y[j].x) < 0) ? 0.f : gscale * (y[j].x - ox),
                                                          (CLAMPED && __float_as_int(y[j].y) < 0) ? 0.f : gscale * (y[j].y - oy));
                }
                if (act && !(fabsf(tot + bs - 1.0f) <= kMassTol)) s_flag[0] = 1;   // NaN-safe
#ifdef CTC_B200_MASSDEV
                if (act) atomicMax(&s_flag[2], __float_as_int(fabsf(tot + bs - 1.0f)));
#endif
                return;
            }
            if constexpr (RC > 1 && YS == 80) {
                // several recursion warps (C3: four), V <= 60, a group of 8 lanes per frame: the RC occupancy rows
                // of the frame with every load issued up front (the looped path below pays one shared-memory
                // latency per load: 28 in a row for RC = 4), the two reductions sharing their shuffle levels
                if (G == 8) {
                    float4 bp[RC];
                    uint2 x[RC][4];
                    float2 y[4];
#pragma unroll
                    for (int rw = 0; rw < RC; ++rw) bp[rw] = *reinterpret_cast<const float4*>(orow + rw * OW + VO + 4 * gl);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = gl + 8 * j;
                        y[j] = make_float2(0.f, 0.f);
#pragma unroll
                        for (int rw = 0; rw < RC; ++rw) x[rw][j] = make_uint2(0u, 0u);
                        if (c < V2) {
#pragma unroll
                            for (int rw = 0; rw < RC; ++rw) x[rw][j] = (reinterpret_cast<const uint2*>(orow + rw * OW))[c];
                            y[j] = y2[c];
                        }
                    }
                    if (act) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int c = gl + 8 * j;
                            if (c < V2) {
#pragma unroll
                                for (int rw = 0; rw < RC; ++rw) (reinterpret_cast<uint2*>(orow + rw * OW))[c] = make_uint2(0u, 0u);
                            }
                        }
                    }
                    float bs = 0.f, tot = 0.f;
#pragma unroll
                    for (int rw = 0; rw < RC; ++rw) bs += (bp[rw].x + bp[rw].y) + (bp[rw].z + bp[rw].w);
                    float2 o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        o[j] = make_float2(0.f, 0.f);
#pragma unroll
                        for (int rw = 0; rw < RC; ++rw) {
                            o[j].x += __uint2float_rn(x[rw][j].x) * (1.0f / kQ31);
                            o[j].y += __uint2float_rn(x[rw][j].y) * (1.0f / kQ31);
                        }
                        tot += o[j].x + o[j].y;
                    }
#pragma unroll
                    for (int sft = 4; sft > 0; sft >>= 1) {
                        bs += __shfl_xor_sync(0xffffffffu, bs, sft);
                        tot += __shfl_xor_sync(0xffffffffu, tot, sft);
                    }
                    if (act) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int c = gl + 8 * j;
                            if (c < V2) {
                                if ((blank >> 1) == c) { if (blank & 1) o[j].y += bs; else o[j].x += bs; }
                                g2[c] = make_float2((CLAMPED && __float_as_int(y[j].x) < 0) ? 0.f : gscale * (y[j].x - o[j].x),
                                                    (CLAMPED && __float_as_int(y[j].y) < 0) ? 0.f : gscale * (y[j].y - o[j].y));
                            }
                        }
                        if (!(fabsf(tot + bs - 1.0f) <= kMassTol)) s_flag[0] = 1;   // NaN-safe
#ifdef CTC_B200_MASSDEV
                        atomicMax(&s_flag[2], __float_as_int(fabsf(tot + bs - 1.0f)));
#endif
                    }
                    return;
                }
            }
            float bs = 0.f;
            for (int i = gl; i < 32 * R; i += G) bs += orow[(i >> 5) * OW + VO + (i & 31)];
            if (!WIDE && !MID && al && R == 1 && V2 <= 4 * G) {     // at most 4 float2 per lane: everything stays in registers
                float2 o[4], y[4];
                float tot = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = gl + j * G;
                    o[j] = make_float2(0.f, 0.f);
                    y[j] = make_float2(0.f, 0.f);
                    if (c < V2) {
                        uint2* p2 = reinterpret_cast<uint2*>(orow) + c;
                        const uint2 x = *p2;
                        y[j] = y2[c];
                        if (act) *p2 = make_uint2(0u, 0u);
                        o[j].x = __uint2float_rn(x.x) * (1.0f / kQ31);
                        o[j].y = __uint2float_rn(x.y) * (1.0f / kQ31);
                        tot += o[j].x + o[j].y;
                    }
                }
                for (int sft = G >> 1; sft > 0; sft >>= 1) {
                    bs += __shfl_xor_sync(0xffffffffu, bs, sft);
                    tot += __shfl_xor_sync(0xffffffffu, tot, sft);
                }
                if (act) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = gl + j * G;
                        if (c < V2) {
                            if ((blank >> 1) == c) { if (blank & 1) o[j].y += bs; else o[j].x += bs; }
                            g2[c] = make_float2((CLAMPED && __float_as_int(y[j].x) < 0) ? 0.f : gscale * (y[j].x - o[j].x),
                                                (CLAMPED && __float_as_int(y[j].y) < 0) ? 0.f : gscale * (y[j].y - o[j].y));
                        }
                    }
                    if (!(fabsf(tot + bs - 1.0f) <= kMassTol)) s_flag[0] = 1;   // NaN-safe
                }
                return;
            }
            if (WIDE && CTC_LIN_WIDE_FAST && V4 <= 8 * 32) {
                // WIDE, a warp per frame (C4): branch-free, explicit 32-bit shared addresses, predicated stores
                const unsigned ya = sbase + (unsigned)lay.y + (unsigned)((int)(ybase - s_y) + fr * Vs) * 4u + (unsigned)gl * 16u;
                const unsigned oa = sbase + (unsigned)lay.occ + (unsigned)((int)(obase - s_occ) + fr * ER) * 4u + (unsigned)gl * 16u;
                float* g4 = reinterpret_cast<float*>(g2) + 4 * gl;
                uint4 x[8];
                float4 y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // (the occupancy rows are the last big region of the CTA's shared memory: no loads beyond the row)
                    x[j] = make_uint4(0u, 0u, 0u, 0u);
                    lds128u_if(oa + (unsigned)(j * 512), x[j], gl + j * 32 < V4);
                    y[j] = lds128(ya + (unsigned)(j * 512));
                    sts128u_if(oa + (unsigned)(j * 512), 0u, 0u, 0u, 0u, act && gl + j * 32 < V4);
                }
                float4 o[8];
                float t0 = 0.f, t1 = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    o[j].x = __uint2float_rn(x[j].x) * (1.0f / kQ31);      // (0 beyond V)
                    o[j].y = __uint2float_rn(x[j].y) * (1.0f / kQ31);
                    o[j].z = __uint2float_rn(x[j].z) * (1.0f / kQ31);
                    o[j].w = __uint2float_rn(x[j].w) * (1.0f / kQ31);
                    t0 += o[j].x + o[j].y;
                    t1 += o[j].z + o[j].w;
                }
                float tot = t0 + t1;
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) {
                    const float b2 = __shfl_xor_sync(0xffffffffu, bs, sft), t2 = __shfl_xor_sync(0xffffffffu, tot, sft);
                    bs += b2;
                    tot += t2;
                }
                const int cb = blank >> 2, kb = blank & 3;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool mine = cb == gl + j * 32;
                    o[j].x += (mine && kb == 0) ? bs : 0.f;
                    o[j].y += (mine && kb == 1) ? bs : 0.f;
                    o[j].z += (mine && kb == 2) ? bs : 0.f;
                    o[j].w += (mine && kb == 3) ? bs : 0.f;
                    stg128_if(g4 + j * 128, (CLAMPED && __float_as_int(y[j].x) < 0) ? 0.f : gscale * (y[j].x - o[j].x),
                              (CLAMPED && __float_as_int(y[j].y) < 0) ? 0.f : gscale * (y[j].y - o[j].y),
                              (CLAMPED && __float_as_int(y[j].z) < 0) ? 0.f : gscale * (y[j].z - o[j].z),
                              (CLAMPED && __float_as_int(y[j].w) < 0) ? 0.f : gscale * (y[j].w - o[j].w), act && gl + j * 32 < V4);
                }
                if (act) {
                    if (!(fabsf(tot + bs - 1.0f) <= kMassTol)) s_flag[0] = 1;   // NaN-safe
#ifdef CTC_B200_MASSDEV
                    atomicMax(&s_flag[2], __float_as_int(fabsf(tot + bs - 1.0f)));
#endif
                }
                return;
            }
            if (!MID && al && YS == 0 && R == 1 && V4 <= 8 * G) {     // wide vocabulary: at most 8 x 128 bit per lane, all loads up front
                uint4* o4 = reinterpret_cast<uint4*>(orow);
                const float4* y4 = reinterpret_cast<const float4*>(y2);
                float4* g4 = reinterpret_cast<float4*>(g2);
                uint4 x[8];
                float4 y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = gl + j * G;
                    x[j] = make_uint4(0u, 0u, 0u, 0u);
                    y[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c < V4) {
                        x[j] = o4[c];
                        y[j] = y4[c];
                        if (act) o4[c] = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                float4 o[8];
                float t0 = 0.f, t1 = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    o[j].x = __uint2float_rn(x[j].x) * (1.0f / kQ31);
                    o[j].y = __uint2float_rn(x[j].y) * (1.0f / kQ31);
                    o[j].z = __uint2float_rn(x[j].z) * (1.0f / kQ31);
                    o[j].w = __uint2float_rn(x[j].w) * (1.0f / kQ31);
                    t0 += o[j].x + o[j].y;
                    t1 += o[j].z + o[j].w;
                }
                float tot = t0 + t1;
                for (int sft = G >> 1; sft > 0; sft >>= 1) {
                    bs += __shfl_xor_sync(0xffffffffu, bs, sft);
                    tot += __shfl_xor_sync(0xffffffffu, tot, sft);
                }
                if (act) {
                    const int cb = blank >> 2, kb = blank & 3;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = gl + j * G;
                        if (c < V4) {
                            const bool mine = cb == c;
                            o[j].x += (mine && kb == 0) ? bs : 0.f;
                            o[j].y += (mine && kb == 1) ? bs : 0.f;
                            o[j].z += (mine && kb == 2) ? bs : 0.f;
                            o[j].w += (mine && kb == 3) ? bs : 0.f;
                            g4[c] = make_float4((CLAMPED && __float_as_int(y[j].x) < 0) ? 0.f : gscale * (y[j].x - o[j].x),
                                                (CLAMPED && __float_as_int(y[j].y) < 0) ? 0.f : gscale * (y[j].y - o[j].y),
                                                (CLAMPED && __float_as_int(y[j].z) < 0) ? 0.f : gscale * (y[j].z - o[j].z),
                                                (CLAMPED && __float_as_int(y[j].w) < 0) ? 0.f : gscale * (y[j].w - o[j].w));
                        }
                    }
                    if (!(fabsf(tot + bs - 1.0f) <= kMassTol)) s_flag[0] = 1;   // NaN-safe
#ifdef CTC_B200_MASSDEV
                    atomicMax(&s_flag[2], __float_as_int(fabsf(tot + bs - 1.0f)));
#endif
                }
                return;
            }
            if (!MID) bs = group_sum(bs, G);      // (MID: reduced together with the class sum, below)
            float tot = 0.f;
            if constexpr (MID) {
                // lane gl holds classes gl, gl + 16, ... of its frame: every load issued up front, immediate offsets
                auto body = [&](auto NVC) {
                    constexpr int NV = decltype(NVC)::value;
                    float* g1 = reinterpret_cast<float*>(g2) + gl;
                    // (explicit 32-bit shared-memory addresses, as in the FIX loops)
                    const unsigned ya = sbase + (unsigned)lay.y + (unsigned)((int)(ybase - s_y) + fr * Vs + gl) * 4u;
                    const unsigned oa = sbase + (unsigned)lay.occ + (unsigned)((int)(obase - s_occ) + fr * ER + gl) * 4u;
                    const int nv = (V - gl + MG - 1) / MG;
                    const int jb = ((blank - gl) & (MG - 1)) == 0 ? (blank - gl) / MG : -1;   // which of my classes is the blank
                    float o[NV], y[NV];
                    float tsum = 0.f;
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        // (branch-free: unconditional loads, a predicated store, selects)
                        const unsigned xq = (unsigned)lds32i(oa + (unsigned)(MG * 4 * j));
                        y[j] = lds32(ya + (unsigned)(MG * 4 * j));
                        sts32i_if(oa + (unsigned)(MG * 4 * j), 0, act && j < nv);
                        o[j] = j < nv ? __uint2float_rn(xq) * (1.0f / kQ31) : 0.f;
                        tsum += o[j];
                    }
                    float bsum = bs;      // the two reductions share their shuffle levels
#pragma unroll
                    for (int sft = MG / 2; sft > 0; sft >>= 1) {
                        const float t2 = __shfl_xor_sync(0xffffffffu, tsum, sft), b2 = __shfl_xor_sync(0xffffffffu, bsum, sft);
                        tsum += t2;
                        bsum += b2;
                    }
                    const float bs = bsum;
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        const float ov = o[j] + (j == jb ? bs : 0.f);
                        stg32_if(g1 + MG * j, (CLAMPED && __float_as_int(y[j]) < 0) ? 0.f : gscale * (y[j] - ov), act && j < nv);
                    }
                    if (act) {
                        if (!(fabsf(tsum + bs - 1.0f) <= kMassTol)) s_flag[0] = 1;   // NaN-safe
#ifdef CTC_B200_MASSDEV
                        atomicMax(&s_flag[2], __float_as_int(fabsf(tsum + bs - 1.0f)));
#endif
                    }
                };
                // (two buckets only: a third one for V <= 64 cost the four-helper launches 5 ... 9 % -- instruction fetch)
                if (V <= 192) body(std::integral_constant<int, 192 / MG>{});
                else body(std::integral_constant<int, 256 / MG>{});
                return;
            }
            if (!WIDE && !al) {      // gradient rows that are not 16-byte aligned in HBM: scalar loads / stores
                float* g1 = reinterpret_cast<float*>(g2);
                const float* y1 = reinterpret_cast<const float*>(y2);
                for (int c = gl; c < V; c += G) {
                    float o = 0.f;
                    for (int rw = 0; rw < R; ++rw) {
                        unsigned* p1 = reinterpret_cast<unsigned*>(orow + rw * OW) + c;
                        o += __uint2float_rn(*p1) * (1.0f / kQ31);
                        if (act) *p1 = 0u;
                    }
                    tot += o;
                    if (c == blank) o += bs;
                    const float y = y1[c];
                    if (act) g1[c] = (CLAMPED && __float_as_int(y) < 0) ? 0.f : gscale * (y - o);
                }
                tot = group_sum(tot, G) + bs;
                if (act && !(fabsf(tot - 1.0f) <= kMassTol)) s_flag[0] = 1;   // NaN-safe
#ifdef CTC_B200_MASSDEV
                if (act) atomicMax(&s_flag[2], __float_as_int(fabsf(tot - 1.0f)));
#endif
                return;
            }
            for (int c = gl; c < V2; c += G) {
                float2 o = make_float2(0.f, 0.f);
                for (int rw = 0; rw < R; ++rw) {
                    uint2* p2 = reinterpret_cast<uint2*>(orow + rw * OW) + c;
                    const uint2 x = *p2;
                    o.x += __uint2float_rn(x.x) * (1.0f / kQ31);
                    o.y += __uint2float_rn(x.y) * (1.0f / kQ31);
                    if (act) *p2 = make_uint2(0u, 0u);
                }
                tot += o.x + o.y;
                if ((blank >> 1) == c) {
                    if (blank & 1) o.y += bs; else o.x += bs;
                }
                const float2 y = y2[c];
                if (act) g2[c] = make_float2((CLAMPED && __float_as_int(y.x) < 0) ? 0.f : gscale * (y.x - o.x),
                                             (CLAMPED && __float_as_int(y.y) < 0) ? 0.f : gscale * (y.y - o.y));
            }
            tot = group_sum(tot, G) + bs;
            if (act && !(fabsf(tot - 1.0f) <= kMassTol)) s_flag[0] = 1;   // NaN-safe
#ifdef CTC_B200_MASSDEV
            if (act) atomicMax(&s_flag[2], __float_as_int(fabsf(tot - 1.0f)));
#endif
        };

        // ---- the helper schedule -----------------------------------------------------------
        // Iteration `it`:  SOFT 0 requests the logits of chunk it + kLinYDist + 1; SOFT: softmax of
        // chunk it; GRAD: gradient rows of chunk it-3 (REC runs chunk it-1, COMB chunk it-2).
        // (WIDE: the logits rows are requested by the first combine warp, which has the most slack -- the first
        // softmax warp spent 536 of its 2756 cycles per chunk issuing two TMA copies)
        const bool iss_acts = !WIDE && ((isA && (ha == 0 || own_rows)) || is_copy);
        const bool do_sm = isA && ha * FA < TC, do_gr = isB && hb * FB < TC && want_grad;
        Ring iss_a(NL), sm_a(NL), gr_a(NL);
        int gr_o = 0;
        int n1h_i = n1, nchh_i = nch;
        asm volatile("" : "+r"(n1h_i), "+r"(nchh_i));
        if (iss_acts) {
            for (int k = 0; k <= kLinYDist; ++k) {    // prologue: logits of chunks 0..kLinYDist
                // (MIDC: a softmax warp copies its frame of chunk 0 itself -- no CTA barrier lies between the prologue
                // and its first softmax -- and the copy warps start with chunk 1)
                const bool mine = !MIDC || (is_copy ? k > 0 : k == 0);
                if (k < nch && mine) issue_logits(k, iss_a.slot);
                else if (cp_groups && mine) cp_async_commit();
                iss_a.advance();
            }
        }
        auto help_iter = [&](int it) {
            LPROF_BEGIN();
            {
                const int ka = it + kLinYDist + 1;
                if (iss_acts && (!MIDC || is_copy)) {
                    if (ka < nch) issue_logits(ka, iss_a.slot);
                    else if (cp_groups) cp_async_commit();   // keep one group per iteration
                }
                iss_a.advance();
            }
            LPROF_SEC(10);
            const int kg = it - 3;
            if (kg >= 0) {
                if (do_gr && kg >= n1h_i && kg < nchh_i && s_flag[1] == 0) {   // gradient rows of chunk it-3
                    int tt0, rows;
                    chunk_at(kg, tt0, rows);
                    if (clp.on) grad_chunk(std::true_type{}, s_occ + (size_t)gr_o * TC * ER, s_y + (size_t)gr_a.slot * TC * Vs, tt0, rows);
                    else grad_chunk(std::false_type{}, s_occ + (size_t)gr_o * TC * ER, s_y + (size_t)gr_a.slot * TC * Vs, tt0, rows);
                }
                if (kg >= n1h_i) gr_o ^= 1;
                gr_a.advance();
            }
            LPROF_SEC(11);
            if (do_sm && it < nchh_i) {               // softmax of chunk `it`
                int tt0, rows;
                chunk_at(it, tt0, rows);
                if (MIDC) { if (it == 0) { cp_async_wait<0>(); __syncwarp(); } }   // (later chunks: the copy warps waited)
                else if (cp_groups) { cp_async_wait<kLinYDist + 1>(); __syncwarp(); }
                else mbar_wait(bar_acts + sm_a.slot, sm_a.parity);
                LPROF_SEC(12);
                if (clp.on) softmax_chunk(std::true_type{}, s_y + (size_t)sm_a.slot * TC * Vs, rows, tt0);
                else softmax_chunk(std::false_type{}, s_y + (size_t)sm_a.slot * TC * Vs, rows, tt0);
            }
            LPROF_SEC(13);
            sm_a.advance();
            // MIDC: the rows of chunk it + 1 have landed before this warp arrives at the barrier that hands them over
            if (is_copy) cp_async_wait<kLinYDist>();
            LPROF_END(it >= n1 + 1);
            __syncthreads();
            if (it == n1) {
                cluster_sync_all();
                __syncthreads();
            }
        };
        int it = 0;
        if constexpr (FIX) {
            // The headline shape class in loops of their own (V = 48, one helper warp, full chunks of 4 frames, a
            // group of 8 lanes per frame), every shared-memory address an offset from the pinned 32-bit base:
            //   phase 1, it in [0, n1 - 1):        logits request for chunk it + 2, softmax of chunk it
            //   phase 2, it in [n1 + 3, nch - 1):  the same plus the gradient rows of chunk it - 3, with the two
            //                                      dependent chains (shuffle trees) interleaved level by level
            // Everything else (short chunks, the phase break, the drain) runs through help_iter.
            static_assert(!FIX || kLinYDist == 1, "the steady-state loops wait for cp.async group it with two groups pending");
            constexpr unsigned YSB = 80u * 4u, YCH = 4u * YSB, ERB = 112u * 4u;
            const unsigned gl8 = (unsigned)(lane & 7) * 8u, fr = (unsigned)(lane >> 3);
            const unsigned y0 = sbase + lay.y + fr * YSB + gl8;          // my 3 x 8 bytes of frame fr, ring slot 0
            const unsigned o0 = sbase + lay.occ + fr * ERB;             // occupancy row of frame fr, buffer 0
            const unsigned fl0 = sbase + lay.flag;
            const ptrdiff_t a_step = (ptrdiff_t)tsign * (ptrdiff_t)frame_stride;   // floats per sweep step
            // logits of chunk ka -> ring slot: two 16-byte cp.async per lane; chunk ka + 6 is pulled into L2
            auto issue_fast = [&](int ka) {
                if (ka < nchh_i) {
                    int tt0, rows;
                    chunk_at(ka, tt0, rows);
                    const float* src = acts_b + (ptrdiff_t)(tbase + tsign * tt0) * (ptrdiff_t)frame_stride;
                    const unsigned dst = sbase + lay.y + (unsigned)iss_a.slot * YCH;
                    if (cp_row[0] < rows) cp_async16_a(dst + (unsigned)cp_dst[0] * 4u, src + cp_src[0]);
                    if (cp_row[1] < rows) cp_async16_a(dst + (unsigned)cp_dst[1] * 4u, src + cp_src[1]);
#if CTC_LIN_PF
                    const int kp = ka + CTC_LIN_PFD;
                    if (kp < nchh_i && (kp < n1h_i) == (ka < n1h_i)) {     // same half of the sweep: 4 * CTC_LIN_PFD steps further on
                        int tp0, rowsp;
                        chunk_at(kp, tp0, rowsp);
#if CTC_LIN_PF == 1
                        if (cp_row[0] < rowsp && pf_lane[0]) prefetch_l2(src + 4 * CTC_LIN_PFD * a_step + cp_src[0]);
                        if (cp_row[1] < rowsp && pf_lane[1]) prefetch_l2(src + 4 * CTC_LIN_PFD * a_step + cp_src[1]);
#else
                        if (lane < rowsp) bulk_prefetch_l2(src + (4 * CTC_LIN_PFD + lane) * a_step, 192u);
#endif
                    }
#endif
                }
                cp_async_commit();
            };
            // softmax of my frame of the chunk in ring slot `ys` (+ gradient row of my frame: occupancy buffer
            // `ob`, softmax rows in ring slot `yg`, gradient row at g2)
            auto help_fast = [&](auto with_grad, auto CL, unsigned ys, unsigned ob, unsigned yg, float2* g2) {
                constexpr bool CLAMPED = decltype(CL)::value;   // compile-time copy of clp.on (uniform branch at the call)
                const Clamp clq{CLAMPED, clp.lo, clp.hi};
                (void)clq;
                constexpr bool WG = decltype(with_grad)::value;
                float4 bp = make_float4(0.f, 0.f, 0.f, 0.f);
                uint2 xo[3];
                float2 yo[3];
                if constexpr (WG) {
                    bp = lds128(ob + 80u * 4u + 2u * gl8);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        xo[j] = lds64u(ob + gl8 + 64u * j);
                        yo[j] = lds64(yg + 64u * j);
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) sts64u(ob + gl8 + 64u * j, make_uint2(0u, 0u));
                }
                cp_async_wait<kLinYDist + 1>();
                __syncwarp();
                float2 x[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) x[j] = lds64(ys + 64u * j);
                unsigned mk = 0u;                  // fused Hardtanh: bit 2j / 2j+1 = gradient blocked
                if (clq.on) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        mk |= (clq.cmask(x[j].x) ? 1u : 0u) << (2 * j) | (clq.cmask(x[j].y) ? 1u : 0u) << (2 * j + 1);
                        x[j].x = clq.cin(x[j].x);
                        x[j].y = clq.cin(x[j].y);
                    }
                }
                if constexpr (VRUN) {              // classes from V on (whatever the ring holds there): -inf, softmax 0
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (2 * ((lane & 7) + 8 * j) >= V) { x[j] = make_float2(-CUDART_INF_F, -CUDART_INF_F); mk &= ~(3u << (2 * j)); }
                }
                float m = fmaxf(fmaxf(fmaxf(x[0].x, x[0].y), fmaxf(x[1].x, x[1].y)), fmaxf(x[2].x, x[2].y));
                float bs = 0.f, tot = 0.f;
                float2 o[3];
                if constexpr (WG) {
                    bs = (bp.x + bp.y) + (bp.z + bp.w);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        o[j].x = __uint2float_rn(xo[j].x) * (1.0f / kQ31);
                        o[j].y = __uint2float_rn(xo[j].y) * (1.0f / kQ31);
                        tot += o[j].x + o[j].y;
                    }
                }
#pragma unroll
                for (int sft = 4; sft > 0; sft >>= 1) {
                    const float m2 = __shfl_xor_sync(0xffffffffu, m, sft);
                    if constexpr (WG) {
                        const float b2 = __shfl_xor_sync(0xffffffffu, bs, sft);
                        const float t2 = __shfl_xor_sync(0xffffffffu, tot, sft);
                        bs += b2;
                        tot += t2;
                    }
                    m = fmaxf(m, m2);
                }
                const float mb = m * kLog2e;
                float z = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    x[j].x = ex2f(fmaf(x[j].x, kLog2e, -mb));
                    x[j].y = ex2f(fmaf(x[j].y, kLog2e, -mb));
                    z += x[j].x + x[j].y;
                }
                if constexpr (WG) {
                    // the blank class sits in column (blank >> 1) of lane (blank >> 1) & 7, slot (blank >> 1) >> 3
                    const int cb = blank >> 1, glq = lane & 7;
                    const float addx = (blank & 1) ? 0.f : bs, addy = (blank & 1) ? bs : 0.f;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const bool mine = cb == glq + 8 * j;
                        const float ox = o[j].x + (mine ? addx : 0.f), oy = o[j].y + (mine ? addy : 0.f);
                        // (a set sign bit of y: the fused Hardtanh blocks this entry's gradient)
                        if (!VRUN || 2 * (glq + 8 * j) < V)
                        g2[glq + 8 * j] = make_float2((CLAMPED && __float_as_int(yo[j].x) < 0) ? 0.f : gscale * (yo[j].x - ox),
                                                      (CLAMPED && __float_as_int(yo[j].y) < 0) ? 0.f : gscale * (yo[j].y - oy));
                    }
                    if (!(fabsf(tot + bs - 1.0f) <= kMassTol)) sts32(fl0, __int_as_float(1));   // NaN-safe
#ifdef CTC_B200_MASSDEV
                    atomicMax(&s_flag[2], __float_as_int(fabsf(tot + bs - 1.0f)));
#endif
                }
                z += __shfl_xor_sync(0xffffffffu, z, 4);
                z += __shfl_xor_sync(0xffffffffu, z, 2);
                z += __shfl_xor_sync(0xffffffffu, z, 1);
                float rs;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(z));
                rs = rs * (2.0f - z * rs);          // one Newton step: full fp32 accuracy
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float a = x[j].x * rs, c = x[j].y * rs;
                    sts64(ys + 64u * j, make_float2((mk >> (2 * j)) & 1u ? -a : a, (mk >> (2 * j + 1)) & 1u ? -c : c));
                }
                if ((lane & 7) == 0) sts32(ys + 48u * 4u, 0.f);     // slot V: what padding pairs gather
            };
            if (iss_acts && do_sm) {     // (always: one helper warp)
                for (; it < n1h_i - 1; ++it) {
                    LPROF_BEGIN();
                    issue_fast(it + 2);
                    iss_a.advance();
                    if (it >= 3) gr_a.advance();
                    if (clp.on) help_fast(std::false_type{}, std::true_type{}, y0 + (unsigned)sm_a.slot * YCH, 0u, 0u, nullptr);
                    else help_fast(std::false_type{}, std::false_type{}, y0 + (unsigned)sm_a.slot * YCH, 0u, 0u, nullptr);
                    sm_a.advance();
                    LPROF_END(false);
                    __syncthreads();
                }
                for (; it < n_it && it < n1h_i + 3; ++it) help_iter(it);
                if (want_grad && it < nchh_i - 1) {
                    // gradient row of my frame of chunk it - 3 (a full chunk of the second half)
                    float* gp = grad_b + (ptrdiff_t)(tbase + tsign * (n_store + (it - 3 - n1h_i) * 4 + (int)fr)) *
                                             (ptrdiff_t)frame_stride;
                    for (; it < nchh_i - 1; ++it) {
                        LPROF_BEGIN();
                        issue_fast(it + 2);
                        iss_a.advance();
                        const unsigned ys = y0 + (unsigned)sm_a.slot * YCH;
                        const unsigned ob = o0 + (unsigned)gr_o * (4u * ERB), yg = y0 + (unsigned)gr_a.slot * YCH;
                        const bool gr = lds32i(fl0 + 4u) == 0;
                        if (clp.on) {
                            if (gr) help_fast(std::true_type{}, std::true_type{}, ys, ob, yg, reinterpret_cast<float2*>(gp));
                            else help_fast(std::false_type{}, std::true_type{}, ys, 0u, 0u, nullptr);
                        } else {
                            if (gr) help_fast(std::true_type{}, std::false_type{}, ys, ob, yg, reinterpret_cast<float2*>(gp));
                            else help_fast(std::false_type{}, std::false_type{}, ys, 0u, 0u, nullptr);
                        }
                        gp += 4 * a_step;
                        gr_o ^= 1;
                        gr_a.advance();
                        sm_a.advance();
                        LPROF_END(true);
                        __syncthreads();
                    }
                }
            }
        }
        for (; it < n_it; ++it) help_iter(it);
    }
#ifdef CTC_B200_PROFILE
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        prof[8] = (unsigned long long)(clock64() - prof_start);
        prof[9] = (unsigned long long)n_it;
    }
#endif
    // every CTA reports whether its half passed; the log-domain kernel redoes flagged utterances
    __syncthreads();
#if defined(CTC_B200_ENDTIME)   // developer build: when did this CTA finish (globaltimer, ns, low bits)
    if (threadIdx.x == 0) {
        unsigned long long t_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        flags[2 * b + (rev ? 1 : 0)] = (int)(t_ & 0x3fffffffull);
    }
#elif defined(CTC_B200_MASSDEV)   // developer build: report the largest posterior-mass deviation instead of the flag
    if (threadIdx.x == 0) flags[2 * b + (rev ? 1 : 0)] = s_flag[1] ? 0x7f800000 : s_flag[2];
#else
    if (threadIdx.x == 0) flags[2 * b + (rev ? 1 : 0)] = s_flag[0];
#endif
    if (queue_ == nullptr) break;
    if (threadIdx.x == 0) {   // the barriers are initialised afresh for the next utterance
        for (int i = 0; i < NL + NS; ++i)
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar_acts + i)) : "memory");
    }
    }   // for (;;): next utterance of the queue
}

}  // namespace ctcb200
