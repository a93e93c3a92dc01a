// ctc_kernels.cuh -- hand-written sm_100a CUDA for the fused CTC loss engine.
//
// Replaces, for jinserk/pytorch-asr's deepspeech_ctc training step, the chain
//   nn.LogSoftmax (asr/models/deepspeech_ctc/network.py:375,395)
//   -> nn.CTCLoss forward (asr/models/trainer.py:153,422 / :508)
//   -> its backward (trainer.py:438 / :517)
// with ONE kernel launch that reads the T x N x V logits and writes nll[N] and
// the gradient with respect to the logits.
//
// Design (see DESIGN.md for the full derivation):
//  * One 2-CTA thread-block cluster per utterance.  CTA rank 0 runs the alpha
//    recursion forward in time, CTA rank 1 runs the beta recursion, expressed as
//    the SAME recursion on the time- and label-reversed problem.  Each CTA
//    stores its lattice rows for the first half of ITS sweep to HBM, the two
//    meet in the middle (hardware cluster barrier, release/acquire), and each
//    then keeps sweeping through the half the partner has already stored,
//    combining its fresh row with the partner's stored row into occupancies and
//    writing the gradient rows as it goes.  The dependent chain is T steps
//    instead of 2T, the parallelism is 2 CTAs per utterance, and only ONE
//    lattice (T x L fp32) ever touches HBM: written once, read once.
//  * The lattice row lives in registers: a thread owns P consecutive (blank,
//    label) cell pairs.  Neighbour exchange is one warp shuffle; the warp-to-warp
//    boundary goes through shared memory.  Rows are stored with 128-bit stores.
//  * Everything is in the base-2 log domain on MUFU.EX2 / MUFU.LG2 (4 per cell
//    pair and step).  Every warp keeps an exact integer offset that it
//    renormalises every 8 steps, so lattice values stay O(100) in magnitude and
//    fp32 rounding does not grow with T (torch's fp32 CTC is ~2e-3 absolute off
//    on the unscaled gradient at T=1000; this kernel ~5e-5).
//  * All global reads (logit rows, partner lattice rows) are cp.async-staged into
//    shared memory one chunk of frames ahead; the log_softmax is fused: a warp
//    per frame, 128-bit accesses, shuffle reductions, log2-probabilities left in
//    shared memory for the label gather.
//  * Cells outside the reachable band (pair i > t, or pair i < S - (T_b - t)) are
//    never computed (warp granularity), stored or loaded; frames t >= T_b cost
//    only the mandatory zero fill of their gradient rows.
#pragma once
#ifndef CTC_LIN_PDL_EARLY
#define CTC_LIN_PDL_EARLY 1
#endif
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <type_traits>

namespace ctcb200 {

constexpr float kNeg = -1.0e30f;        // finite stand-in for log(0); absorbing under +
constexpr float kRealThresh = -1.0e29f; // anything below is "log(0)"
constexpr float kLog2e = 1.4426950408889634f;
constexpr double kLn2 = 0.6931471805599453;
constexpr int kRenormMask = 7;          // renormalise warp offsets every 8 steps
constexpr int kMaxChunk = 8;            // frames per staged chunk

struct FusedParams {
    const float* __restrict__ acts;        // [T, N, V] logits
    const int32_t* __restrict__ targets;   // concatenated labels
    const int32_t* __restrict__ tgt_off;   // [N] start of each utterance's labels
    const int32_t* __restrict__ in_lens;   // [N]
    const int32_t* __restrict__ tgt_lens;  // [N]
    const float* __restrict__ grad_scale;  // [N] or nullptr
    float* __restrict__ nll;               // [N]
    float* __restrict__ grad;              // [T, N, V] or nullptr
    float* __restrict__ lattice;           // workspace: [n_utt][T][row_stride]
    float* __restrict__ ysave;             // workspace: [n_utt][T][V] softmax rows (ctc_lin.cuh) or nullptr
    int* __restrict__ status;              // device status word (bit flags)
    long long lat_utt_stride;              // floats per utterance in `lattice`
    int T, N, V, blank, zero_infinity;
    int utt_begin;                         // first utterance handled by this launch
    int row_stride;                        // floats per lattice row = 2*NP + Wp
    int chunk;                             // frames per chunk (<= kMaxChunk)
    // Layout of acts / grad: element (t, b, v) lives at t * frame_stride + b * utt_stride + v.
    //   time-major  [T,N,V] (trainer.py:418):  frame_stride = N*V, utt_stride = V
    //   batch-major [N,T,V] (what network.py:380-396 emits; folds trainer.py:418's transpose+copy):
    //                                          frame_stride = V,   utt_stride = T*V
    long long frame_stride, utt_stride;
    // Fused Hardtanh(clamp_lo, clamp_hi) in front of the log_softmax (network.py:370); its backward
    // mask (gradient 0 where the raw logit is outside the open interval) is applied to the gradient.
    int use_clamp;
    float clamp_lo, clamp_hi;
    const int* __restrict__ redo;          // [2 * N] or nullptr: run only utterances the linear kernel flagged
};

enum StatusBits : int {
    kStatusBadLabel = 1,     // a target label outside [0, V)
    kStatusBadLength = 2,    // input length outside [0, T] or target length < 0 / too long
    kStatusPeerTimeout = 4,  // fused loss all-reduce: a peer's pair did not arrive in time
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// log2(2^a + 2^b) for finite a, b (kNeg stands for -inf and is absorbing).
__device__ __forceinline__ float lse2(float a, float b) {
    return fmaxf(a, b) + lg2f(1.0f + ex2f(-fabsf(a - b)));
}
// Fused Hardtanh (network.py:370) in front of the log_softmax.  `cin` is the value that enters the
// softmax; `cmask` says whether Hardtanh's backward blocks the gradient (raw logit outside the OPEN
// interval, torch's hardtanh_backward).  The log-domain kernels keep the mask in the lowest mantissa bit
// of the stored log2-probability (1 ulp, only when the clamp is on); the linear kernel in the sign bit
// of the stored probability.
struct Clamp {
    bool on;
    float lo, hi;
    __device__ __forceinline__ float cin(float x) const { return on ? fminf(fmaxf(x, lo), hi) : x; }
    __device__ __forceinline__ bool cmask(float x) const { return on && !(x > lo && x < hi); }
    __device__ __forceinline__ float tag(float lp, float raw) const {
        if (!on) return lp;
        return __int_as_float((__float_as_int(lp) & ~1) | (cmask(raw) ? 1 : 0));
    }
    __device__ __forceinline__ bool tagged(float lp) const { return on && (__float_as_int(lp) & 1); }
};

__device__ __forceinline__ float warp_max(float x) {
    float m;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(x));
    return m;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\t"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Deterministic block-wide reductions through shared memory (fixed order).
__device__ __forceinline__ float block_max(float x, float* s_red, int w, int lane, int W) {
    x = warp_max(x);
    if (lane == 0) s_red[w] = x;
    __syncthreads();
    float m = s_red[0];
    for (int i = 1; i < W; ++i) m = fmaxf(m, s_red[i]);
    __syncthreads();
    return m;
}
__device__ __forceinline__ float block_sum(float x, float* s_red, int w, int lane, int W) {
    x = warp_sum(x);
    if (lane == 0) s_red[w] = x;
    __syncthreads();
    float s = s_red[0];
    for (int i = 1; i < W; ++i) s += s_red[i];
    __syncthreads();
    return s;
}

// Shared-memory carve-up, shared by host (size) and device (pointers).
//   lp2:   2 x chunk rows of Vs floats: raw logits staged by cp.async, turned into
//          log2-probabilities in place; slot V of every row holds kNeg (the value an
//          out-of-range label gathers)
//   stage: 2 x chunk partner lattice rows
struct SmemLayout {
    int lab, cstart, lp2, eB, eY, stage, bnd, red, total;  // byte offsets
    int Vs;                                                // floats per lp2 row
    __host__ __device__ static int up(int x, int a) { return (x + a - 1) / a * a; }
    __host__ __device__ SmemLayout(int NP, int W, int V, int chunk, int row_stride) {
        Vs = up(V + 1, 4);
        int o = 0;
        lab = o;    o += up(NP * 4, 16);
        cstart = o; o += up((V + 2) * 4, 16);
        lp2 = o;    o += up(2 * chunk * Vs * 4, 16);
        eB = o;     o += up(chunk * NP * 4, 16);
        eY = o;     o += up(chunk * NP * 4, 16);
        stage = o;  o += up(2 * chunk * row_stride * 4, 16);
        bnd = o;    o += up(2 * (W + 1) * 8, 16);
        red = o;    o += up(32 * 4, 16);
        total = o;
    }
};

template <int P> struct VecOf;
template <> struct VecOf<1> { using type = float; };
template <> struct VecOf<2> { using type = float2; };
template <> struct VecOf<4> { using type = float4; };

template <int P>
__device__ __forceinline__ void store_vec(float* dst, const float (&v)[P]) {
    if constexpr (P == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    else if constexpr (P == 2) *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
    else {
#pragma unroll
        for (int k = 0; k < P; ++k) dst[k] = v[k];
    }
}

// ---------------------------------------------------------------------------
// The fused kernel.  P = lattice pairs (blank cell + label cell) per thread.
// grid = 2 * n_utt CTAs in clusters of 2; block = NT threads (multiple of 32);
// NP = NT * P >= S_max + 1.
// ---------------------------------------------------------------------------
template <int P, int MAXT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MAXT, MAXT <= 256 ? 2 : 1)
ctc_fused_kernel(const FusedParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int NT = blockDim.x, W = NT >> 5, NP = NT * P;
    const int b = p.utt_begin + (blockIdx.x >> 1);
    const bool rev = (blockIdx.x & 1) != 0;
    // fallback mode (vocabularies the log-domain pipe kernel cannot take, V % 4 != 0): both CTAs of the
    // cluster leave unless the linear kernel flagged the utterance
    if (p.redo != nullptr) {
#if CTC_LIN_PDL_EARLY
        asm volatile("griddepcontrol.launch_dependents;");
#endif
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if ((p.redo[2 * b] | p.redo[2 * b + 1]) == 0) return;
    }
    const int T = p.T, V = p.V, blank = p.blank;
    const int RS = p.row_stride, TC = p.chunk;
    const bool vec4 = (V & 3) == 0 && ((p.frame_stride | p.utt_stride) & 3) == 0;
    const Clamp clp{p.use_clamp != 0, p.clamp_lo, p.clamp_hi};

    const SmemLayout lay(NP, W, V, TC, RS);
    const int Vs = lay.Vs;
    int* s_lab = reinterpret_cast<int*>(smem_raw + lay.lab);
    int* s_cstart = reinterpret_cast<int*>(smem_raw + lay.cstart);
    float* s_lp2 = reinterpret_cast<float*>(smem_raw + lay.lp2);
    float* s_eB = reinterpret_cast<float*>(smem_raw + lay.eB);
    float* s_eY = reinterpret_cast<float*>(smem_raw + lay.eY);
    float* s_stage = reinterpret_cast<float*>(smem_raw + lay.stage);
    float2* s_bnd = reinterpret_cast<float2*>(smem_raw + lay.bnd);
    float* s_red = reinterpret_cast<float*>(smem_raw + lay.red);

    int Tb = p.in_lens[b], S = p.tgt_lens[b];
    if (Tb < 0 || Tb > T || S < 0 || S > NP - 1) {
        if (tid == 0) atomicOr(p.status, kStatusBadLength);
        Tb = min(max(Tb, 0), T);
        S = min(max(S, 0), NP - 1);
    }
    const int32_t* tg = p.targets + p.tgt_off[b];
    const bool want_grad = p.grad != nullptr;
    const float gscale = p.grad_scale ? p.grad_scale[b] : 1.0f;
    const size_t frame_stride = (size_t)p.frame_stride;      // floats between frames
    const float* acts_b = p.acts + (size_t)b * (size_t)p.utt_stride;
    float* grad_b = want_grad ? p.grad + (size_t)b * (size_t)p.utt_stride : nullptr;

    // ---- mandatory zero fill of gradient rows t >= T_b (no compute) ----------
    if (want_grad) {
        const int nrows = T - Tb;
        const int mine = (nrows + (rev ? 0 : 1)) >> 1;  // rows Tb+rev, Tb+rev+2, ...
        const int first = Tb + (rev ? 1 : 0);
        if (vec4) {
            const int V4 = V >> 2;
            for (int r = w; r < mine; r += W) {
                float4* g4 = reinterpret_cast<float4*>(grad_b + (size_t)(first + 2 * r) * frame_stride);
                for (int c = lane; c < V4; c += 32) g4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (int r = w; r < mine; r += W) {
                float* g = grad_b + (size_t)(first + 2 * r) * frame_stride;
                for (int c = lane; c < V; c += 32) g[c] = 0.f;
            }
        }
    }
    if (Tb == 0) {  // torch: empty input => 0 for an empty target, +inf otherwise
        if (!rev && tid == 0)
            p.nll[b] = (S == 0 || p.zero_infinity) ? 0.0f : CUDART_INF_F;
        return;  // both CTAs of the cluster take this exit
    }

    // ---- per-utterance setup: labels in sweep order, class-sorted positions ---
    for (int i = tid; i < NP; i += NT) {
        int c = V;  // out-of-range pairs gather the kNeg slot of the lp2 row
        if (i < S) {
            c = rev ? tg[S - 1 - i] : tg[i];
            if (c < 0 || c >= V) {
                atomicOr(p.status, kStatusBadLabel);
                c = min(max(c, 0), V - 1);
            }
        }
        s_lab[i] = c;
    }
    for (int v = tid; v < V + 2; v += NT) s_cstart[v] = 0;
    for (int i = tid; i < 2 * (W + 1); i += NT) s_bnd[i] = make_float2(kNeg, 0.f);
    __syncthreads();

    const int i0 = tid * P;  // first pair of this thread
    int lab[P], pos[P];
    bool skip[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const int i = i0 + k;
        lab[k] = s_lab[i];
        skip[k] = (i >= 1 && i < S && lab[k] != s_lab[i - 1]);
        pos[k] = 0;
        if (want_grad && i < S) atomicAdd(&s_cstart[lab[k] + 1], 1);
    }
    __syncthreads();
    if (want_grad) {
        // exclusive scan of the class histogram (warp 0, 32 classes per round)
        if (w == 0) {
            int carry = 0;
            for (int base = 0; base < V + 1; base += 32) {
                const int v = base + lane;
                int inc = (v < V + 1) ? s_cstart[v] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += y;
                }
                if (v < V + 1) s_cstart[v] = carry + inc;  // = #labels of class < v
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        __syncthreads();
        // deterministic rank inside the class: equal labels keep sweep order
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const int i = i0 + k;
            if (i < S) {
                int r = 0;
                for (int j = 0; j < i; ++j) r += (s_lab[j] == lab[k]) ? 1 : 0;
                pos[k] = s_cstart[lab[k]] + r;
            }
        }
    }

    // ---- sweep geometry --------------------------------------------------------
    // Pair i is inside the band at sweep step tt iff 0 <= tt - i <= C (label cell) /
    // <= C - 1 (blank cell), C = T_b - S: reachable from the start and able to finish.
    const int C = max(Tb - S, -1);
    unsigned winB[P], winY[P];  // (unsigned)(tt - i) < win  <=>  cell in band (0 for padding pairs)
#pragma unroll
    for (int k = 0; k < P; ++k) {
        winB[k] = (i0 + k <= S) ? (unsigned)max(C, 0) : 0u;
        winY[k] = (i0 + k < S) ? (unsigned)(C + 1) : 0u;
    }
    const unsigned win_thread = (i0 <= S) ? (unsigned)(C + P) : 0u;   // any cell of the thread
    const int w_first = (32 * P * w <= S) ? 32 * P * w : 0x3fffffff;  // warp active for tt in
    const int w_last = C + 32 * P * w + 32 * P - 1;                   //   [w_first, w_last]
    const int Tm = Tb >> 1;
    const int n_store = rev ? (Tb - Tm) : Tm;  // rows this CTA stores; the rest it consumes
    float* lat_b = p.lattice + (size_t)(b - p.utt_begin) * (size_t)p.lat_utt_stride;
    const int tsign = rev ? -1 : 1, tbase = rev ? Tb - 1 : 0;  // frame of sweep step tt = tbase + tsign*tt

    // partner's view of my cells: my blank i is its blank S-i, my label i its label S-1-i
    const int jB0 = S - i0;                          // partner pair index of my first blank
    const int pw_hi = max(jB0, 0) / (32 * P);        // at most two partner warps per thread
    const int pw_lo = max(pw_hi - 1, 0);
    bool selB[P], selY[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        selB[k] = max(jB0 - k, 0) / (32 * P) == pw_hi;
        selY[k] = max(jB0 - 1 - k, 0) / (32 * P) == pw_hi;
    }

    // chunk schedule: n1 store chunks, then n2 consume chunks
    const int n1 = (n_store + TC - 1) / TC;
    const int n2 = want_grad ? (Tb - n_store + TC - 1) / TC : (Tb > n_store ? 1 : 0);
    const int nch = n1 + n2;
    auto chunk_at = [&](int c, int& tt0, int& rows) {
        if (c < n1) { tt0 = c * TC; rows = min(TC, n_store - tt0); }
        else { tt0 = n_store + (c - n1) * TC; rows = want_grad ? min(TC, Tb - tt0) : 1; }
    };
    // cp.async staging of the logit rows of a chunk into lp2 buffer `ab`
    auto issue_acts = [&](int ab, int tt0, int rows) {
        float* dst0 = s_lp2 + (size_t)ab * TC * Vs;
        for (int r = w; r < rows; r += W) {
            const int t = tbase + tsign * (tt0 + r);
            const float* src = acts_b + (size_t)t * frame_stride;
            float* dst = dst0 + r * Vs;
            if (vec4) for (int c = lane * 4; c < V; c += 128) cp_async16(dst + c, src + c);
            else for (int c = lane; c < V; c += 32) cp_async4(dst + c, src + c);
        }
    };
    // cp.async staging of the partner's lattice rows of a consume chunk (band-limited)
    auto issue_partner = [&](int pb, int tt0, int rows) {
        float* dst0 = s_stage + (size_t)pb * TC * RS;
        for (int r = w; r < rows; r += W) {
            const int tt = tt0 + r;
            const int t = tbase + tsign * tt;
            // my in-band pairs [max(S-(Tb-tt),0), min(tt,S)] <-> partner pairs S - that, and S-1 - that
            const int mlo = max(S - (Tb - tt), 0), mhi = min(tt, S);
            const int plo = max(S - 1 - mhi, 0), phi = S - mlo;
            const float* src = lat_b + (size_t)t * RS;
            float* dst = dst0 + (size_t)r * RS;
            const int c0 = plo & ~3;
            for (int c = c0 + lane * 4; c <= phi; c += 128) {
                cp_async16(dst + c, src + c);
                cp_async16(dst + NP + c, src + NP + c);
            }
            for (int c = 2 * NP + lane * 4; c < RS; c += 128) cp_async16(dst + c, src + c);
        }
    };

    // ---- sweep state ----------------------------------------------------------
    float aB[P], aY[P];
#pragma unroll
    for (int k = 0; k < P; ++k) { aB[k] = kNeg; aY[k] = kNeg; }
    if (tid == 0) aB[0] = 0.0f;  // virtual row "-1": log(1) in front of the first blank
    float off = 0.0f;            // this warp's exact integer offset (true = off + a)
    bool fresh = (w != 0);       // warp has not received any real value yet
    int par = 0;
    float ll_int = 0.f, ll_frac = 0.f;
    bool infeasible = false;

    // ---- one recursion step; FIRST = the first consumed row (computes the likelihood)
    auto step = [&](auto consume_tag, auto first_tag, int tt, int r, const float* lp2,
                    const float* st) {
        constexpr bool CONSUME = decltype(consume_tag)::value;
        constexpr bool FIRST = decltype(first_tag)::value;
        const bool wact = (tt >= w_first) && (tt <= w_last);
        float eBv[P], eYv[P], nBv[P], nYv[P];
        if (CONSUME) {
#pragma unroll
            for (int k = 0; k < P; ++k) { eBv[k] = kNeg; eYv[k] = kNeg; nBv[k] = 0.f; nYv[k] = 0.f; }
        }
        if (wact) {
            const float lpb = lp2[blank];
            float am1 = __shfl_up_sync(0xffffffffu, aY[P - 1], 1);
            float2 bq;
            if (lane == 0) bq = s_bnd[par * (W + 1) + w];   // slot 0 is the constant (kNeg, 0)
            if (fresh) {
                const float bv = __shfl_sync(0xffffffffu, bq.x, 0);
                const float bo = __shfl_sync(0xffffffffu, bq.y, 0);
                if (bv > kRealThresh) { off = bo; fresh = false; }
            }
            if (lane == 0) am1 = bq.x + (bq.y - off);
            float lpl[P];
#pragma unroll
            for (int k = 0; k < P; ++k) {
                lpl[k] = lp2[lab[k]];
                const float x = lse2(aB[k], am1);
                const float yin = skip[k] ? x : aB[k];
                const float ynew = lpl[k] + lse2(aY[k], yin);
                am1 = aY[k];
                aB[k] = lpb + x;
                aY[k] = ynew;
            }
            // exact integer renormalisation keeps |a| small (fp32 accuracy at long T)
            if ((tt & kRenormMask) == kRenormMask) {
                float m = kNeg;
#pragma unroll
                for (int k = 0; k < P; ++k) m = fmaxf(m, fmaxf(aB[k], aY[k]));
                m = warp_max(m);
                if (m > kRealThresh) {
                    const float sh = rintf(m);
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        aB[k] = fmaxf(aB[k] - sh, kNeg);
                        aY[k] = fmaxf(aY[k] - sh, kNeg);
                    }
                    off += sh;
                    fresh = false;
                }
            }
            if (lane == 31) s_bnd[(par ^ 1) * (W + 1) + w + 1] = make_float2(aY[P - 1], off);

            const unsigned u0 = (unsigned)(tt - i0);
            if (!CONSUME) {
                if (want_grad || tt == n_store - 1) {
                    float* row = lat_b + (size_t)(tbase + tsign * tt) * RS;
                    if (u0 < win_thread) {
                        store_vec<P>(row + i0, aB);
                        store_vec<P>(row + NP + i0, aY);
                    }
                    if (lane == 0) row[2 * NP + w] = off;
                }
            } else {
                const float ohi = off + st[2 * NP + pw_hi], olo = off + st[2 * NP + pw_lo];
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const unsigned u = u0 - (unsigned)k;
                    if (u < winB[k]) {
                        eBv[k] = (aB[k] + st[jB0 - k]) - lpb;
                        nBv[k] = selB[k] ? ohi : olo;
                    }
                    if (u < winY[k]) {
                        eYv[k] = (aY[k] + st[NP + jB0 - 1 - k]) - lpl[k];
                        nYv[k] = selY[k] ? ohi : olo;
                    }
                }
            }
        }
        if (CONSUME) {
            if (FIRST) {
                // log-likelihood from the first combined row: ll = ll_int + ll_frac
                float pm = kNeg;
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    if (eBv[k] > kRealThresh) pm = fmaxf(pm, nBv[k] + rintf(eBv[k]));
                    if (eYv[k] > kRealThresh) pm = fmaxf(pm, nYv[k] + rintf(eYv[k]));
                }
                pm = block_max(pm, s_red, w, lane, W);
                float z = 0.f;
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    if (eBv[k] > kRealThresh) z += ex2f((nBv[k] - pm) + eBv[k]);
                    if (eYv[k] > kRealThresh) z += ex2f((nYv[k] - pm) + eYv[k]);
                }
                z = block_sum(z, s_red, w, lane, W);
                infeasible = !(pm > kRealThresh);
                ll_int = infeasible ? 0.f : pm;
                ll_frac = infeasible ? 0.f : lg2f(z);
                if (!rev && tid == 0) {
                    float out;
                    if (infeasible) out = p.zero_infinity ? 0.0f : CUDART_INF_F;
                    else out = (float)(-((double)ll_int + (double)ll_frac) * kLn2);
                    p.nll[b] = out;
                }
            }
            if (want_grad) {
                float eBo[P];
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    eBo[k] = (eBv[k] - ll_frac) + (nBv[k] - ll_int);   // kNeg stays ~kNeg
                    if (i0 + k < S) s_eY[r * NP + pos[k]] = (eYv[k] - ll_frac) + (nYv[k] - ll_int);
                }
                store_vec<P>(s_eB + r * NP + i0, eBo);
            }
        }
        par ^= 1;
        __syncthreads();
    };

    using TrueT = std::integral_constant<bool, true>;
    using FalseT = std::integral_constant<bool, false>;

    // ---- main loop over chunks ---------------------------------------------------
    int ab = 0, pb = 0;
    {
        int tt0, rows;
        chunk_at(0, tt0, rows);
        issue_acts(0, tt0, rows);
        cp_async_commit();
    }
    for (int c = 0; c < nch; ++c) {
        int tt0, rows;
        chunk_at(c, tt0, rows);
        const bool consume = c >= n1;
        if (c == n1) {
            // the partner has stored every row I am about to consume (and vice versa)
            cluster_sync_all();
            issue_partner(pb, tt0, rows);
            cp_async_commit();
        }
        if (c + 1 < nch) {  // prefetch everything chunk c+1 reads from global memory
            int ntt0, nrows;
            chunk_at(c + 1, ntt0, nrows);
            issue_acts(ab ^ 1, ntt0, nrows);
            if (c + 1 > n1) issue_partner(pb ^ 1, ntt0, nrows);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        // ---- fused log_softmax, in place on the staged rows: one warp per frame --
        float* lp2c = s_lp2 + (size_t)ab * TC * Vs;
        for (int r = w; r < rows; r += W) {
            float* row = lp2c + r * Vs;
            float m = -CUDART_INF_F, z = 0.f;
            if (vec4) {
                float4* row4 = reinterpret_cast<float4*>(row);
                const int V4 = V >> 2;
                if (V4 <= 32) {  // whole row in one float4 per lane: single pass
                    float4 q = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
                    if (lane < V4) q = row4[lane];
                    const float4 raw = q;
                    if (lane < V4) { q.x = clp.cin(q.x); q.y = clp.cin(q.y); q.z = clp.cin(q.z); q.w = clp.cin(q.w); }
                    m = warp_max(fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
                    q.x = (q.x - m) * kLog2e; q.y = (q.y - m) * kLog2e;
                    q.z = (q.z - m) * kLog2e; q.w = (q.w - m) * kLog2e;
                    z = warp_sum(ex2f(q.x) + ex2f(q.y) + ex2f(q.z) + ex2f(q.w));
                    const float lz = lg2f(z);
                    if (lane < V4)
                        row4[lane] = make_float4(clp.tag(fmaxf(q.x - lz, kNeg), raw.x), clp.tag(fmaxf(q.y - lz, kNeg), raw.y),
                                                 clp.tag(fmaxf(q.z - lz, kNeg), raw.z), clp.tag(fmaxf(q.w - lz, kNeg), raw.w));
                } else {
                    for (int c4 = lane; c4 < V4; c4 += 32) {
                        const float4 q = row4[c4];
                        m = fmaxf(m, fmaxf(fmaxf(clp.cin(q.x), clp.cin(q.y)), fmaxf(clp.cin(q.z), clp.cin(q.w))));
                    }
                    m = warp_max(m);
                    for (int c4 = lane; c4 < V4; c4 += 32) {
                        const float4 q = row4[c4];
                        z += ex2f((clp.cin(q.x) - m) * kLog2e) + ex2f((clp.cin(q.y) - m) * kLog2e) +
                             ex2f((clp.cin(q.z) - m) * kLog2e) + ex2f((clp.cin(q.w) - m) * kLog2e);
                    }
                    z = warp_sum(z);
                    const float lz = lg2f(z);
                    for (int c4 = lane; c4 < V4; c4 += 32) {
                        const float4 q = row4[c4];
                        row4[c4] = make_float4(clp.tag(fmaxf((clp.cin(q.x) - m) * kLog2e - lz, kNeg), q.x),
                                               clp.tag(fmaxf((clp.cin(q.y) - m) * kLog2e - lz, kNeg), q.y),
                                               clp.tag(fmaxf((clp.cin(q.z) - m) * kLog2e - lz, kNeg), q.z),
                                               clp.tag(fmaxf((clp.cin(q.w) - m) * kLog2e - lz, kNeg), q.w));
                    }
                }
            } else {
                for (int v = lane; v < V; v += 32) m = fmaxf(m, clp.cin(row[v]));
                m = warp_max(m);
                for (int v = lane; v < V; v += 32) z += ex2f((clp.cin(row[v]) - m) * kLog2e);
                z = warp_sum(z);
                const float lz = lg2f(z);
                for (int v = lane; v < V; v += 32) {
                    const float raw = row[v];
                    row[v] = clp.tag(fmaxf((clp.cin(raw) - m) * kLog2e - lz, kNeg), raw);
                }
            }
            if (lane == 0) row[V] = kNeg;  // what padding pairs gather
        }
        __syncthreads();

        // ---- lattice recursion over the chunk ---------------------------------
        const float* stc = s_stage + (size_t)pb * TC * RS;
        if (!consume) {
            for (int r = 0; r < rows; ++r)
                step(FalseT{}, FalseT{}, tt0 + r, r, lp2c + r * Vs, stc);
        } else {
            int r = 0;
            if (c == n1) { step(TrueT{}, TrueT{}, tt0, 0, lp2c, stc); r = 1; }
            if (want_grad)
                for (; r < rows; ++r)
                    step(TrueT{}, FalseT{}, tt0 + r, r, lp2c + r * Vs, stc + (size_t)r * RS);
        }

        // ---- gradient rows of the chunk: one warp per frame ---------------------
        if (consume && want_grad) {
            for (int r = w; r < rows; r += W) {
                const int t = tbase + tsign * (tt0 + r);
                float* g = grad_b + (size_t)t * frame_stride;
                const float* lp2 = lp2c + r * Vs;
                const float* eB = s_eB + r * NP;
                const float* eY = s_eY + r * NP;
                if (infeasible) {
                    const float fill = p.zero_infinity ? 0.0f : CUDART_NAN_F;
                    for (int v = lane; v < V; v += 32) g[v] = fill;
                    continue;
                }
                float bs = 0.f;
                for (int i = lane; i <= S; i += 32) bs += ex2f(eB[i]);
                bs = warp_sum(bs);
                for (int v = lane; v < V; v += 32) {
                    float occ = (v == blank) ? bs : 0.f;
                    const int k1 = s_cstart[v + 1];
                    for (int k = s_cstart[v]; k < k1; ++k) occ += ex2f(eY[k]);
                    const float lpv = lp2[v];
                    g[v] = clp.tagged(lpv) ? 0.0f : gscale * (ex2f(lpv) - occ);
                }
            }
            pb ^= 1;
        }
        ab ^= 1;
        __syncthreads();
    }
    if (n_store == Tb) cluster_sync_all();  // T_b == 1: the beta CTA has nothing to consume
}

}  // namespace ctcb200
