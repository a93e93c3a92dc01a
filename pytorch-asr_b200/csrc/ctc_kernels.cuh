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
//  * The lattice row lives in registers: thread i owns the blank cell 2i and the
//    label cell 2i+1.  Neighbour exchange is one warp shuffle; the warp-to-warp
//    boundary goes through shared memory.
//  * Everything is in the base-2 log domain on MUFU.EX2 / MUFU.LG2.  Every warp
//    keeps an exact integer offset that it renormalises every 8 steps, so lattice
//    values stay O(100) in magnitude and fp32 rounding does not grow with T
//    (torch's fp32 CTC loses ~2e-3 absolute on the unscaled gradient at T=1000).
//  * The log_softmax is fused: a warp per frame reads the V logits with 128-bit
//    loads, reduces with shuffles and leaves log2-probabilities in shared memory.
//  * Cells outside the reachable band (s > 2t+1 or s < L-2(T_b-t)) are never
//    computed (warp granularity), stored or loaded; frames t >= T_b cost only the
//    mandatory zero fill of their gradient rows.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace ctcb200 {

constexpr float kNeg = -1.0e30f;        // finite stand-in for log(0)
constexpr float kRealThresh = -1.0e29f; // anything below is "log(0)"
constexpr float kLog2e = 1.4426950408889634f;
constexpr double kLn2 = 0.6931471805599453;
constexpr int kRenormMask = 7;          // renormalise warp offsets every 8 steps
constexpr int kMaxChunk = 8;            // frames per softmax/gradient chunk

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
    int* __restrict__ status;              // device status word (bit flags)
    long long lat_utt_stride;              // floats per utterance in `lattice`
    int T, N, V, blank, zero_infinity;
    int utt_begin;                         // first utterance handled by this launch
    int row_stride;                        // floats per lattice row = 2*NP + Wp
    int chunk;                             // frames per chunk (<= kMaxChunk)
};

enum StatusBits : int {
    kStatusBadLabel = 1,     // a target label outside [0, V)
    kStatusBadLength = 2,    // input length outside [0, T] or target length < 0 / too long
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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

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
struct SmemLayout {
    int lab, cstart, lp2, eB, eY, stage, bnd, red, total;  // byte offsets
    __host__ __device__ static int up(int x, int a) { return (x + a - 1) / a * a; }
    __host__ __device__ SmemLayout(int NP, int W, int V, int chunk, int row_stride) {
        int o = 0;
        lab = o;    o += up(NP * 4, 16);
        cstart = o; o += up((V + 2) * 4, 16);
        lp2 = o;    o += up(chunk * up(V, 4) * 4, 16);
        eB = o;     o += up(chunk * NP * 4, 16);
        eY = o;     o += up(chunk * NP * 4, 16);
        stage = o;  o += up(2 * chunk * row_stride * 4, 16);
        bnd = o;    o += up(2 * W * 8, 16);
        red = o;    o += up(32 * 4, 16);
        total = o;
    }
};

// ---------------------------------------------------------------------------
// The fused kernel.  P = lattice pairs (blank cell + label cell) per thread.
// grid = 2 * n_utt CTAs in clusters of 2; block = NT threads (multiple of 32);
// NP = NT * P >= S_max + 1.
// ---------------------------------------------------------------------------
template <int P>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(1024, 1)
ctc_fused_kernel(const FusedParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int NT = blockDim.x, W = NT >> 5, NP = NT * P;
    const int b = p.utt_begin + (blockIdx.x >> 1);
    const bool rev = (blockIdx.x & 1) != 0;
    const int T = p.T, N = p.N, V = p.V, blank = p.blank;
    const int Vp = (V + 3) & ~3;
    const int RS = p.row_stride, TC = p.chunk;
    const bool vec4 = (V & 3) == 0;

    const SmemLayout lay(NP, W, V, TC, RS);
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
    const int L = 2 * S + 1;
    const int32_t* tg = p.targets + p.tgt_off[b];
    const bool want_grad = p.grad != nullptr;
    const float gscale = p.grad_scale ? p.grad_scale[b] : 1.0f;

    // ---- mandatory zero fill of gradient rows t >= T_b (no compute) ----------
    if (want_grad) {
        const int nrows = T - Tb;
        const int mine = (nrows + (rev ? 0 : 1)) >> 1;  // rows Tb+rev, Tb+rev+2, ...
        if (vec4) {
            const int V4 = V >> 2;
            for (int idx = tid; idx < mine * V4; idx += NT) {
                int r = idx / V4, c = idx - r * V4;
                int t = Tb + (rev ? 1 : 0) + 2 * r;
                reinterpret_cast<float4*>(p.grad + ((size_t)t * N + b) * V)[c] =
                    make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (int idx = tid; idx < mine * V; idx += NT) {
                int r = idx / V, c = idx - r * V;
                int t = Tb + (rev ? 1 : 0) + 2 * r;
                p.grad[((size_t)t * N + b) * V + c] = 0.f;
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
        int c = -1;
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
    for (int i = tid; i < 2 * W; i += NT) s_bnd[i] = make_float2(kNeg, 0.f);
    __syncthreads();

    int lab[P], pos[P];
    bool skip[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const int i = tid * P + k;
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
                int v = base + lane;
                int x = (v < V + 1) ? s_cstart[v] : 0;
                int inc = x;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += y;
                }
                if (v < V + 1) s_cstart[v] = carry + inc;  // inclusive of bucket v-1 counts
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        __syncthreads();
        // s_cstart[v] now = number of labels with class < v  (start of class v)
        // deterministic rank inside the class: labels of equal class keep sweep order
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const int i = tid * P + k;
            if (i < S) {
                int r = 0;
                for (int j = 0; j < i; ++j) r += (s_lab[j] == lab[k]) ? 1 : 0;
                pos[k] = s_cstart[lab[k]] + r;
            }
        }
    }

    // ---- sweep state ----------------------------------------------------------
    float aB[P], aY[P];
#pragma unroll
    for (int k = 0; k < P; ++k) { aB[k] = kNeg; aY[k] = kNeg; }
    if (tid == 0) aB[0] = 0.0f;  // virtual row "-1": log(1) in front of the first blank
    float off = 0.0f;            // this warp's exact integer offset (true = off + a)
    bool fresh = (w != 0);       // warp has not received any real value yet
    int par = 0;
    float ll_int = 0.f, ll_frac = kNeg;
    bool infeasible = false;

    const int Tm = Tb >> 1;
    const int n_store = rev ? (Tb - Tm) : Tm;  // rows this CTA stores; the rest it consumes
    float* lat_b = p.lattice + (size_t)(b - p.utt_begin) * (size_t)p.lat_utt_stride;
    const int wc_lo = 2 * P * 32 * w;          // first / last lattice cell of this warp
    const int wc_hi = wc_lo + 2 * P * 32 - 1;
    const int nchunk4 = RS >> 2;

    // stage partner lattice rows [tt0, tt0+rows) into s_stage[buf] (cp.async, band-limited)
    auto issue_stage = [&](int buf, int tt0, int rows) {
        for (int idx = tid; idx < rows * nchunk4; idx += NT) {
            const int r = idx / nchunk4, c = idx - r * nchunk4;
            const int tt = tt0 + r;
            const int t = rev ? (Tb - 1 - tt) : tt;
            // my in-band cells [lo_c, hi_c] <-> partner cells [L-1-hi_c, L-1-lo_c]
            const int lo_c = max(L - 2 * (Tb - tt), 0), hi_c = min(2 * tt + 1, L - 1);
            const int pp_lo = max((L - 1 - hi_c) >> 1, 1) - 1;  // partner pair range (conservative)
            const int pp_hi = (L - 1 - lo_c) >> 1;
            const int f = c << 2;                                // float index inside the row
            bool need;
            if (f >= 2 * NP) need = true;                        // warp offsets
            else {
                const int q = (f >= NP) ? f - NP : f;            // pair index of the first float
                need = (q + 3 >= pp_lo) && (q <= pp_hi);
            }
            if (need)
                cp_async16(s_stage + ((size_t)(buf * TC + r)) * RS + f,
                           lat_b + (size_t)t * RS + f);
        }
        cp_async_commit();
    };

    int buf = 0;
    for (int tt0 = 0; tt0 < Tb;) {
        const bool consume = tt0 >= n_store;
        if (tt0 == n_store) {
            // Partner has stored every row I am about to consume (and vice versa).
            cluster_sync_all();
            // forward only: the first consumed row yields the likelihood, nothing else is read
            issue_stage(buf, tt0, want_grad ? min(TC, Tb - tt0) : 1);
        }
        const int rows = min(TC, (consume ? Tb : n_store) - tt0);

        if (consume) cp_async_wait_all();

        // ---- fused log_softmax: one warp per frame, 128-bit loads ------------
        for (int r = w; r < rows; r += W) {
            const int tt = tt0 + r;
            const int t = rev ? (Tb - 1 - tt) : tt;
            const float* x = p.acts + ((size_t)t * N + b) * V;
            float* out = s_lp2 + r * Vp;
            float m = -CUDART_INF_F, z = 0.f;
            if (vec4) {
                const float4* x4 = reinterpret_cast<const float4*>(x);
                const int V4 = V >> 2;
                for (int c = lane; c < V4; c += 32) {
                    float4 q = __ldg(x4 + c);
                    m = fmaxf(m, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
                }
                m = warp_max(m);
                for (int c = lane; c < V4; c += 32) {
                    float4 q = __ldg(x4 + c);
                    z += ex2f((q.x - m) * kLog2e) + ex2f((q.y - m) * kLog2e) +
                         ex2f((q.z - m) * kLog2e) + ex2f((q.w - m) * kLog2e);
                }
                z = warp_sum(z);
                const float lz = lg2f(z);
                for (int c = lane; c < V4; c += 32) {
                    float4 q = __ldg(x4 + c);
                    float4 o;
                    o.x = fmaxf((q.x - m) * kLog2e - lz, kNeg);
                    o.y = fmaxf((q.y - m) * kLog2e - lz, kNeg);
                    o.z = fmaxf((q.z - m) * kLog2e - lz, kNeg);
                    o.w = fmaxf((q.w - m) * kLog2e - lz, kNeg);
                    reinterpret_cast<float4*>(out)[c] = o;
                }
            } else {
                for (int v = lane; v < V; v += 32) m = fmaxf(m, __ldg(x + v));
                m = warp_max(m);
                for (int v = lane; v < V; v += 32) z += ex2f((__ldg(x + v) - m) * kLog2e);
                z = warp_sum(z);
                const float lz = lg2f(z);
                for (int v = lane; v < V; v += 32)
                    out[v] = fmaxf((__ldg(x + v) - m) * kLog2e - lz, kNeg);
            }
        }
        // prefetch the partner rows of the NEXT chunk while this one is processed
        if (consume && want_grad && tt0 + rows < Tb) issue_stage(buf ^ 1, tt0 + rows, min(TC, Tb - tt0 - rows));
        __syncthreads();

        // ---- lattice recursion over the chunk ---------------------------------
        for (int r = 0; r < rows; ++r) {
            const int tt = tt0 + r;
            const int t = rev ? (Tb - 1 - tt) : tt;
            const int lo_c = L - 2 * (Tb - tt), hi_c = 2 * tt + 1;
            const bool wact = (wc_lo <= hi_c) && (wc_hi >= lo_c) && (wc_lo < L);
            const float* lp2 = s_lp2 + r * Vp;
            float vB[P], vY[P], nB[P], nY[P];
#pragma unroll
            for (int k = 0; k < P; ++k) { vB[k] = kNeg; vY[k] = kNeg; nB[k] = 0.f; nY[k] = 0.f; }

            if (wact) {
                const float lpb = lp2[blank];
                // left neighbour's label cell from the previous row
                float am1 = __shfl_up_sync(0xffffffffu, aY[P - 1], 1);
                float2 bq = make_float2(kNeg, 0.f);
                if (lane == 0 && w > 0) bq = s_bnd[par * W + (w - 1)];
                if (fresh) {
                    const float bv = __shfl_sync(0xffffffffu, bq.x, 0);
                    const float bo = __shfl_sync(0xffffffffu, bq.y, 0);
                    if (bv > kRealThresh) { off = bo; fresh = false; }
                }
                if (lane == 0) am1 = (w > 0) ? bq.x + (bq.y - off) : kNeg;

                float lpl[P];
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const int i = tid * P + k;
                    lpl[k] = (i < S) ? lp2[lab[k]] : kNeg;
                    const float x = lse2(aB[k], am1);
                    const float yin = skip[k] ? x : aB[k];
                    const float ynew = lpl[k] + lse2(aY[k], yin);
                    const float bnew = lpb + x;
                    am1 = aY[k];
                    aB[k] = (i <= S) ? bnew : kNeg;
                    aY[k] = (i < S) ? ynew : kNeg;
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
                if (lane == 31) s_bnd[(par ^ 1) * W + w] = make_float2(aY[P - 1], off);

                if (!consume) {
                    if (want_grad || tt == n_store - 1) {
                        float* row = lat_b + (size_t)t * RS;
#pragma unroll
                        for (int k = 0; k < P; ++k) {
                            const int i = tid * P + k;
                            const int sB = 2 * i, sY = 2 * i + 1;
                            if (i <= S && sB >= lo_c && sB <= hi_c) row[i] = aB[k];
                            if (i < S && sY >= lo_c && sY <= hi_c) row[NP + i] = aY[k];
                        }
                        if (lane == 0) row[2 * NP + w] = off;
                    }
                } else {
                    const float* st = s_stage + ((size_t)(buf * TC + r)) * RS;
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        const int i = tid * P + k;
                        const int sB = 2 * i, sY = 2 * i + 1;
                        if (i <= S && sB >= lo_c && sB <= hi_c) {
                            const int j = S - i;  // partner's pair index of this blank cell
                            vB[k] = aB[k] + st[j] - lpb;
                            nB[k] = off + st[2 * NP + j / (32 * P)];
                        }
                        if (i < S && sY >= lo_c && sY <= hi_c) {
                            const int j = S - 1 - i;
                            vY[k] = aY[k] + st[NP + j] - lpl[k];
                            nY[k] = off + st[2 * NP + j / (32 * P)];
                        }
                    }
                }
            }

            if (consume) {
                if (tt == n_store) {
                    // log-likelihood from the first combined row: ll = ll_int + ll_frac
                    float pm = kNeg;
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        if (vB[k] > kRealThresh) pm = fmaxf(pm, nB[k] + rintf(vB[k]));
                        if (vY[k] > kRealThresh) pm = fmaxf(pm, nY[k] + rintf(vY[k]));
                    }
                    pm = block_max(pm, s_red, w, lane, W);
                    float z = 0.f;
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        if (vB[k] > kRealThresh) z += ex2f((nB[k] - pm) + vB[k]);
                        if (vY[k] > kRealThresh) z += ex2f((nY[k] - pm) + vY[k]);
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
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        const int i = tid * P + k;
                        s_eB[r * NP + i] = (vB[k] > kRealThresh) ? (vB[k] - ll_frac) + (nB[k] - ll_int) : kNeg;
                        if (i < S)
                            s_eY[r * NP + pos[k]] = (vY[k] > kRealThresh) ? (vY[k] - ll_frac) + (nY[k] - ll_int) : kNeg;
                    }
                }
            }
            par ^= 1;
            __syncthreads();
            if (consume && !want_grad) break;  // forward only: nll is known
        }
        if (consume && !want_grad) break;

        // ---- gradient rows of the chunk: one warp per frame ---------------------
        if (consume) {
            for (int r = w; r < rows; r += W) {
                const int tt = tt0 + r;
                const int t = rev ? (Tb - 1 - tt) : tt;
                float* g = p.grad + ((size_t)t * N + b) * V;
                const float* lp2 = s_lp2 + r * Vp;
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
                    g[v] = gscale * (ex2f(lp2[v]) - occ);
                }
            }
            buf ^= 1;
        }
        __syncthreads();
        tt0 += rows;
    }
    if (n_store == Tb) cluster_sync_all();  // T_b == 1: the beta CTA has nothing to consume
}

// ---------------------------------------------------------------------------
// grad[t, b, :] *= scale[b] (or *= scale[0] when `per_utt` is 0), skipped
// entirely -- no memory traffic -- for factors equal to 1.  This is the whole
// "backward": the fused kernel already wrote d nll / d logits; autograd only
// has to apply the upstream grad_output (trainer.py:429 loss.mul_(0), AMP loss
// scaling at :435-436), which is exactly 1 in the plain fp32 step.
// ---------------------------------------------------------------------------
__global__ void ctc_scale_grad_kernel(float* __restrict__ grad, const float* __restrict__ scale,
                                      int per_utt, int T, int N, int V) {
    const int b = blockIdx.x;
    const float s = per_utt ? scale[b] : scale[0];
    if (s == 1.0f) return;
    const size_t total = (size_t)T * V;
    for (size_t idx = (size_t)blockIdx.y * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.y * blockDim.x) {
        const size_t t = idx / V, v = idx - t * V;
        float* g = grad + (t * N + b) * V + v;
        *g = (s == 0.0f) ? 0.0f : *g * s;
    }
}

// ---------------------------------------------------------------------------
// out[0] = sum_b nll_b / max(S_b, 1)   (mode 1, 'mean' numerator; trainer.py:153)
//        = sum_b nll_b                 (mode 2, 'sum')
// out[1] = N  (the normaliser that is all-reduced together with out[0])
// Single CTA, fixed summation order => bit-reproducible.
// ---------------------------------------------------------------------------
__global__ void ctc_reduce_loss_kernel(const float* __restrict__ nll,
                                       const int32_t* __restrict__ tgt_lens, int N, int mode,
                                       float* __restrict__ out, float* __restrict__ loss) {
    __shared__ double s_part[32];
    double acc = 0.0;
    for (int b = threadIdx.x; b < N; b += blockDim.x) {
        double v = (double)nll[b];
        if (mode == 1) v /= (double)max(tgt_lens[b], 1);
        acc += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += s_part[i];
        out[0] = (float)s;
        out[1] = (float)N;
        if (loss) loss[0] = (mode == 1) ? (float)(s / (double)max(N, 1)) : (float)s;
    }
}

}  // namespace ctcb200
