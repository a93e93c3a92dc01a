// ctc_lin.cuh -- the LINEAR-domain warp-specialised fused CTC kernel for sm_100a.
//
// Same decomposition as ctc_pipe.cuh (a 2-CTA cluster per utterance, the alpha CTA and the
// time/label-reversed beta CTA meet in the middle, ONE fp32 lattice goes through HBM, REC /
// HELP warp roles, TMA-staged rings), but the lattice recursion runs on PROBABILITIES, not on
// log-probabilities:
//
//      x      = aB + aY_prev                       (blank cell, before the emission)
//      inner  = aY + (skip ? x : aB)               (label cell, before the emission)
//      aB'    = y_blank * x ;  aY' = y_label * inner
//
// i.e. 5 FP32 instructions per cell pair and step and NO MUFU (the log-domain kernel spends
// 4 MUFU + ~12 FP32 per pair).  The occupancy pass needs no exp either: occupancy =
// alpha * beta~ / P with beta~ the partner's PRE-emission value, so nothing is divided by y.
//
// Range: every THREAD keeps an exact power-of-two exponent `off` for its P pairs (true value =
// a * 2^off).  Neighbour values are brought to the receiver's scale with one exact multiply;
// a thread renormalises its cells to ~2^32 at every chunk end; an incoming value far above the
// receiver's scale makes the receiver rescale first (warp-uniform rare branch).  Rows are
// stored with their per-thread exponents ([blank plane][label plane][exponents]).
//
// Safety net: whatever the scaling loses (a cell flushed to zero, a clamped scale) shows up
// as missing posterior mass.  The helpers check  |sum_s occupancy_t(s) - 1| <= kMassTol  for
// EVERY frame; an utterance that fails (or whose likelihood underflows to 0, which includes
// every infeasible utterance) is flagged in `flags` and recomputed by the log-domain kernel
// (ctc_pipe_kernel, launched right after with the same grid; clusters of unflagged utterances
// exit at once).  The linear path is therefore exact to fp32 rounding or not used at all.
//
// Alignment trick: the alpha CTA shifts its lattice by delta = (P-1-S) mod P slots, so that the
// P partner cells a consumer thread needs are exactly ONE partner thread's P cells, reversed:
// 128-bit conflict-free shared-memory loads and a single partner exponent per thread.
#pragma once
#include "ctc_pipe.cuh"

namespace ctcb200 {

constexpr int kLinTarget = 32;          // a thread's largest cell is renormalised to ~2^32
constexpr int kLinK = 40;               // a thread's scale is at most 2^40 below the scale of the thread under it
constexpr int kLinFresh = -(1 << 24);   // exponent of a thread that has not received anything yet
constexpr int kLinHmax = 44;            // clamp of the combine exponent (no overflow of p * 2^h)
constexpr float kMassTol = 3.0e-5f;     // |sum of occupancies - 1| per frame
constexpr int kLinNone = -(1 << 28);    // exponent of a term that is exactly zero

__device__ __forceinline__ int clamp_exp(int e) { return max(min(e, 127), -127); }
// 2^e for e in [-126, 127]; 0 for e <= -127 (flush); 2^127 above
__device__ __forceinline__ float pow2c(int e) { return __int_as_float((clamp_exp(e) + 127) << 23); }
// exponent of x >= 0 (zero / denormals give -127)
__device__ __forceinline__ int expo(float x) { return (__float_as_int(x) >> 23) - 127; }
__device__ __forceinline__ int warp_max_i(int x) { return __reduce_max_sync(0xffffffffu, x); }

template <int P>
__device__ __forceinline__ void store_row(float* dst, const float (&v)[P]) {
    if constexpr (P == 8) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
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
__device__ __forceinline__ void load_row(const float* src, float (&v)[P]) {
    if constexpr (P == 8) {
        const float4 a = *reinterpret_cast<const float4*>(src);
        const float4 b = *reinterpret_cast<const float4*>(src + 4);
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

// lattice row = [blank plane: NP][label plane: NP][per-thread exponents: NP / P]
__host__ __device__ __forceinline__ int lin_row_stride(int NP, int P) { return 2 * NP + (NP / P + 3) / 4 * 4; }

// Template parameters: P pairs per thread; RC = number of recursion warps when it is known at
// compile time (1: the common case S + P <= 32 * P, every stride becomes an immediate) or 0 for
// "run time"; YS = floats per row of the emission ring (64 for V <= 60) or 0 for "run time".
template <int P, int RC, int YS, int MAXT, int MINB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MAXT, MINB)
ctc_lin_kernel(const PipeParams pp, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FusedParams& p = pp.f;
    const int NT = blockDim.x, NW = NT >> 5;
    const int R = RC > 0 ? RC : pp.R, H = pp.H, NP = RC > 0 ? 32 * P * RC : pp.NP;
    const int lane = threadIdx.x & 31;
    const int shift = pp.rotate > 0 ? (int)((blockIdx.x / pp.rotate) * R) % NW : 0;
    const int w = ((int)(threadIdx.x >> 5) + NW - shift) % NW;   // role (virtual) warp id
    const int tid = w * 32 + lane;
    const int b = p.utt_begin + (blockIdx.x >> 1);
    const bool rev = (blockIdx.x & 1) != 0;
    const int T = p.T, N = p.N, V = p.V, blank = p.blank;
    const int RS = RC > 0 ? lin_row_stride(32 * P * RC, P) : p.row_stride, TC = p.chunk, D = pp.D;
    const bool is_rec = w < R;
    const int hw = w - R;  // helper index (>= 0 for helpers)

    const PipeSmem lay(NP, R, V, TC, RS, D, YS);
    const int Vs = YS > 0 ? YS : lay.Vs, ER = NP + ypad(NP) + 4, NL = lay.NL, NS = lay.NS;
    int* s_lab = reinterpret_cast<int*>(smem_raw + lay.lab);
    int* s_pos = reinterpret_cast<int*>(smem_raw + lay.pos);
    int* s_cstart = reinterpret_cast<int*>(smem_raw + lay.cstart);
    int* s_fill = reinterpret_cast<int*>(smem_raw + lay.fill);
    float* s_y = reinterpret_cast<float*>(smem_raw + lay.lp2);      // emission probabilities ring
    float* s_e = reinterpret_cast<float*>(smem_raw + lay.e);
    float* s_stage = reinterpret_cast<float*>(smem_raw + lay.stage);
    float2* s_bnd = reinterpret_cast<float2*>(smem_raw + lay.bnd);
    float* s_red = reinterpret_cast<float*>(smem_raw + lay.red);
    int* s_flag = reinterpret_cast<int*>(smem_raw + lay.ll);        // [0] redo, [1] no gradient rows
    uint64_t* bar_acts = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);   // [NL]
    uint64_t* bar_part = bar_acts + NL;                                      // [NS]

    int Tb = p.in_lens[b], S = p.tgt_lens[b];
    if (Tb < 0 || Tb > T || S < 0 || S > NP - P) {
        if (tid == 0) atomicOr(p.status, kStatusBadLength);
        Tb = min(max(Tb, 0), T);
        S = min(max(S, 0), NP - P);
    }
    const int32_t* tg = p.targets + p.tgt_off[b];
    const bool want_grad = p.grad != nullptr;
    const float gscale = p.grad_scale ? p.grad_scale[b] : 1.0f;
    const size_t frame_stride = (size_t)N * V;
    const float* acts_b = p.acts + (size_t)b * V;
    float* grad_b = want_grad ? p.grad + (size_t)b * V : nullptr;
    const int V4 = V >> 2;

    // ---- helpers: mandatory zero fill of gradient rows t >= T_b (no compute) --------
    if (want_grad && !is_rec) {
        const int nrows = T - Tb;
        const int mine = (nrows + (rev ? 0 : 1)) >> 1;  // rows Tb+rev, Tb+rev+2, ...
        float* g = grad_b + (size_t)(Tb + (rev ? 1 : 0) + 2 * hw) * frame_stride;
        const size_t ginc = 2 * (size_t)H * frame_stride;
        for (int r = hw; r < mine; r += H, g += ginc)
            for (int c = lane; c < V4; c += 32) reinterpret_cast<float4*>(g)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (Tb == 0) {  // torch: empty input => 0 for an empty target, +inf otherwise
        if (tid == 0) {
            flags[2 * b + (rev ? 1 : 0)] = 0;
            if (!rev) p.nll[b] = (S == 0 || p.zero_infinity) ? 0.0f : CUDART_INF_F;
        }
        return;  // both CTAs of the cluster take this exit
    }

    // ---- per-utterance setup (all warps) --------------------------------------------
    // slot s of this CTA holds pair i = s - delta.  SS = S + delta == P-1 (mod P) for the alpha
    // CTA's delta; the beta CTA uses delta = 0.  My slot s <-> partner blank slot SS - s, partner
    // label slot SS - 1 - s.
    const int SS = S + ((P - 1 - S) & (P - 1));
    const int delta = rev ? 0 : SS - S;
    for (int s = tid; s < NP; s += NT) {
        const int i = s - delta;
        int c = V;  // padding pairs gather the zero slot of the y row
        int ps = (i < 0) ? S + s : s;   // padding label cells park their zeros behind the sorted ones
        if (i >= 0 && i < S) {
            c = rev ? tg[S - 1 - i] : tg[i];
            if (c < 0 || c >= V) {
                atomicOr(p.status, kStatusBadLabel);
                c = min(max(c, 0), V - 1);
            }
        }
        s_lab[s] = c;
        s_pos[s] = ps;
    }
    for (int v = tid; v < V + 2; v += NT) s_cstart[v] = 0;
    for (int i = tid; i < 2 * (R + 1); i += NT) s_bnd[i] = make_float2(0.f, 0.f);
    if (tid == 0) {
        s_flag[0] = 0; s_flag[1] = 0;
        for (int i = 0; i < NL; ++i) mbar_init(bar_acts + i, 32);   // 32 lanes' cp.async
        for (int i = 0; i < NS; ++i) mbar_init(bar_part + i, 1);    // one TMA producer
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();
    if (want_grad)
        for (int i = tid; i < S; i += NT) atomicAdd(&s_cstart[s_lab[i + delta] + 1], 1);
    __syncthreads();
    if (want_grad && w == 0) {
        // exclusive scan of the class histogram: s_cstart[v] = #labels of class < v
        int carry = 0;
        for (int base = 0; base < V + 1; base += 32) {
            const int v = base + lane;
            int inc = (v < V + 1) ? s_cstart[v] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            if (v < V + 1) { s_cstart[v] = carry + inc; s_fill[v] = carry + inc; }
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        __syncwarp();
        // class-sorted position of every label, equal labels in sweep order (deterministic)
        for (int base = 0; base < S; base += 32) {
            const int i = base + lane;
            const int c = (i < S) ? s_lab[i + delta] : -1 - lane;
            const unsigned peers = __match_any_sync(0xffffffffu, c);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            int first = 0;
            if (i < S) first = s_fill[c];
            __syncwarp();
            if (i < S) {
                s_pos[i + delta] = first + rank;
                if (rank == 0) s_fill[c] = first + __popc(peers);
            }
            __syncwarp();
        }
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
    auto chunk_at = [&](int c, int& tt0, int& rows) {  // first sweep step / row count of chunk c
        if (c < n1) { tt0 = c * TC; rows = min(TC, n_store - tt0); }
        else { tt0 = n_store + (c - n1) * TC; rows = want_grad ? min(TC, Tb - tt0) : 1; }
    };

    if (is_rec) {
        // =============================================================================
        // REC: lattice recursion on probabilities
        // =============================================================================
        const int s0 = tid * P;          // my first slot
        const int i0 = s0 - delta;       // my first pair (may be negative: leading padding)
        float skf[P];                    // 1 if label cell k also takes the skip transition, else 0
        const float* yk[P];              // &y[label k] in row 0 of the current chunk of the emission ring
        float* ek[P];                    // class-sorted position of label cell k in row 0 of the e chunk
        int lab[P], ysl[P];
        bool vB[P], vY[P];
        float aB[P], aY[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const int i = i0 + k;
            lab[k] = s_lab[s0 + k];
            skf[k] = (i >= 1 && i < S && lab[k] != s_lab[s0 + k - 1]) ? 1.0f : 0.0f;
            ysl[k] = NP + ypad(s_pos[s0 + k]);   // where my label cell goes in an e row
            vB[k] = i >= 0 && i <= S;
            vY[k] = i >= 0 && i < S;
            aB[k] = 0.f;
            aY[k] = 0.f;
        }
        int off = kLinFresh;             // my exponent: true value = a * 2^off
        if (i0 <= 0 && i0 + P > 0) {     // the thread that owns pair 0: virtual row "-1" = 1
#pragma unroll
            for (int k = 0; k < P; ++k) if (i0 + k == 0) aB[k] = 1.0f;
            off = 0;
        }
        // partner thread of my P cells (blank k <-> its blank P-1-k; label k <-> its label P-2-k,
        // my last label <-> the last label of the thread below it)
        const int M = (SS - (P - 1)) / P;
        const int X = M - tid;
        const bool hasX = X >= 0 && X < 32 * R, hasX1 = X >= 1 && X - 1 < 32 * R;
        int E0 = 0;                      // integer part of log2 P(labels | logits)
        float rz = 0.f;                  // 1 / mantissa sum: occupancy = a * p~ * 2^(off+o-E0) * rz
        // thread / warp activity windows in sweep steps (widened by P pairs, see ctc_pipe.cuh)
        unsigned win_store = (i0 - P <= S) ? (unsigned)(C + 3 * P) : 0u;   // (u + P) < win
        unsigned win_cons = (i0 <= S && hasX) ? (unsigned)(C + P) : 0u;    // u < win
        // per-step constants live in registers (the compiler would otherwise re-derive them from the
        // kernel parameters inside the unrolled steps)
        int wg_i = want_grad ? 1 : 0, c_i = C;
        const int offd = 2 * NP + tid - s0;   // my exponent's position relative to my blank vector
        asm volatile("" : "+r"(win_store), "+r"(win_cons), "+r"(wg_i), "+r"(c_i));
        const int iw = 32 * P * w - delta;   // first pair of my warp
        const int w_first = (iw - P <= S) ? iw - P : 0x3fffffff;
        const int w_last = C + iw + 32 * P - 1 + P;
        float2* bnd_rd = s_bnd + w;              // slot 0 is the constant (0, 0)
        float2* bnd_wr = s_bnd + (R + 1) + w + 1;
        const bool lane0 = lane == 0, lane31 = (lane == 31) && R > 1;
        const int nbar = 32 * R;
        const bool wg = wg_i != 0;
        const int row_step = tsign * RS;
        int u = -i0;                     // sweep step minus my first pair: tt - i0

        // Renormalisation, every min(P, 4) steps, off the dependent chain.  Every thread wants its
        // largest cell at 2^kLinTarget; in addition no thread may sit more than 2^kLinK below the
        // thread under it (a prefix maximum over the lanes of  exponent + K * thread), so that a value
        // handed up by the neighbour never overflows the receiver's scale: the per-step exchange
        // then needs neither a test nor a branch.  Threads whose pairs can all no longer finish are
        // cleared; threads nothing has reached yet inherit a scale from below.
        auto renorm = [&]() {
            float m = 0.f;
#pragma unroll
            for (int k = 0; k < P; ++k) m = fmaxf(m, fmaxf(aB[k], aY[k]));
            const bool gone = u - (P - 1) > c_i + 1;          // even my last pair is dead
            const bool live = m > 0.f && !gone;
            int h = (live ? off + expo(m) : kLinFresh) + kLinK * tid;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, h, o);
                if (lane >= o) h = max(h, y);
            }
            if constexpr (RC != 1) {
                if (R > 1) {                                   // carry the prefix maximum across warps
                    int* red = reinterpret_cast<int*>(s_red);
                    if (lane == 31) red[w] = h;
                    named_bar_sync(1, nbar);
                    for (int i = 0; i < w; ++i) h = max(h, red[i]);
                    named_bar_sync(1, nbar);
                }
            }
            const int g = h - kLinK * tid;
            const int noff = g < kLinFresh / 2 ? kLinFresh : g - kLinTarget;
            const int sh = live ? max(noff - off, -126) : 0;   // cells *= 2^-sh
            float f = sh > 126 ? 0.f : __int_as_float((127 - sh) << 23);
            if (gone) f = 0.f;
#pragma unroll
            for (int k = 0; k < P; ++k) { aB[k] *= f; aY[k] *= f; }
            off = live ? off + sh : noff;
        };
        // one recursion step: a[t] <- a[t-1]; xs / ins = the cells BEFORE the emission (scale `off`);
        // yo = float offset of this step's row inside the chunk of the emission ring
        auto advance = [&](const float* yb_ptr, int yo, float (&xs)[P], float (&ins)[P]) {
            const float yb = yb_ptr[yo];
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
                const float yl = yk[k][yo];
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

        Ring ring_y(NL), ring_part(NS);   // position of the chunk REC works on
        int e_buf = 0;
        renorm();                         // hands every thread above pair 0 its initial scale
        for (int it = 0; it < nch + 2; ++it) {
            const int k = it - 1;
            if (k >= 0 && k < nch) {
                int tt0, rows;
                chunk_at(k, tt0, rows);
                const float* ychunk = s_y + (size_t)ring_y.slot * TC * Vs;
                const float* yb_ptr = ychunk + blank;
#pragma unroll
                for (int q = 0; q < P; ++q) yk[q] = ychunk + lab[q];
                if (k < n1) {
                    // ---- store chunk: pre-emission rows go to HBM for the partner ----------
                    // one running offset from the lattice base; everything else is an immediate
                    long long roff = (long long)(b - p.utt_begin) * p.lat_utt_stride +
                                     (long long)(tbase + tsign * tt0) * RS + s0;
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
                                store_row<P>(row, xs);
                                store_row<P>(row + NP, ins);
                                *reinterpret_cast<int*>(row + offd) = off;
                            }
                        }
                        row += row_step;
                        end_step();
                        if (P < 4 && (r + 1) % P == 0 && r + 1 < rows) renorm();
                    }
                    renorm();
                } else {
                    // ---- consume chunk: combine with the partner's stored rows ----------
                    mbar_wait(bar_part + ring_part.slot, ring_part.parity);   // TMA data landed
                    const float* st = s_stage + ((size_t)ring_part.slot * TC + (rev ? rows - 1 : 0)) * RS;
                    float* erow = s_e + (size_t)e_buf * TC * ER;
#pragma unroll
                    for (int q = 0; q < P; ++q) ek[q] = erow + ysl[q];
                    float* eb = erow + s0;
                    const float* stp = st + (hasX ? X * P : 0);         // partner thread's P blanks
                    const int* sto = reinterpret_cast<const int*>(st + 2 * NP) + (hasX ? X : 0);
                    float pb[P], py[P];    // partner cells matching my blank k / label k
                    int ob, oy;            // partner exponents: of thread X and of thread X-1
                    auto fetch = [&]() {
                        float qb[P], qy[P];
                        load_row<P>(stp, qb);
                        load_row<P>(stp + NP, qy);
                        ob = *sto;
#pragma unroll
                        for (int q = 0; q < P; ++q) pb[q] = qb[P - 1 - q];
#pragma unroll
                        for (int q = 0; q + 1 < P; ++q) py[q] = qy[P - 2 - q];
                        // my last label pairs with the LAST label of partner thread X-1 = the thread
                        // lane+1 of my warp talks to; lane 31 reads it itself
                        float yl = __shfl_down_sync(0xffffffffu, qy[P - 1], 1);
                        int ol = __shfl_down_sync(0xffffffffu, ob, 1);
                        if (lane == 31 && hasX1) { yl = stp[NP - 1]; ol = sto[-1]; }
                        py[P - 1] = yl;
                        oy = ol;
                    };
                    int r0 = 0;
                    if (k == n1) {
                        // first combined row: also yields the likelihood P = sum_s a * p~.
                        // Exponent/mantissa form: E0 = max exponent of any term, z = sum of the
                        // terms scaled by 2^-E0; log2 P = E0 + log2 z.
                        float tB[P], tY[P];
                        int eB[P], eY[P];
#pragma unroll
                        for (int q = 0; q < P; ++q) { tB[q] = 0.f; tY[q] = 0.f; eB[q] = kLinNone; eY[q] = kLinNone; }
                        if (warp_on(tt0)) {
                            float xs[P], ins[P];
                            fetch();
                            advance(yb_ptr, 0, xs, ins);
                            if ((unsigned)u < win_cons) {
#pragma unroll
                                for (int q = 0; q < P; ++q) {
                                    if (vB[q] && aB[q] > 0.f && pb[q] > 0.f) {
                                        const int ea = expo(aB[q]), ep = expo(pb[q]);
                                        tB[q] = (aB[q] * pow2c(-ea)) * (pb[q] * pow2c(-ep));
                                        eB[q] = ea + ep + off + ob;
                                    }
                                    const bool okY = q + 1 < P ? true : hasX1;
                                    const int oq = q + 1 < P ? ob : oy;
                                    if (vY[q] && okY && aY[q] > 0.f && py[q] > 0.f) {
                                        const int ea = expo(aY[q]), ep = expo(py[q]);
                                        tY[q] = (aY[q] * pow2c(-ea)) * (py[q] * pow2c(-ep));
                                        eY[q] = ea + ep + off + oq;
                                    }
                                }
                            }
                        }
                        int em = kLinNone;
#pragma unroll
                        for (int q = 0; q < P; ++q) em = max(em, max(eB[q], eY[q]));
                        em = warp_max_i(em);
                        if (R > 1) {
                            if (lane0) reinterpret_cast<int*>(s_red)[w] = em;
                            named_bar_sync(1, nbar);
                            for (int i = 0; i < R; ++i) em = max(em, reinterpret_cast<int*>(s_red)[i]);
                            named_bar_sync(1, nbar);
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
                            if (lane0) s_red[w] = z;
                            named_bar_sync(1, nbar);
                            z = 0.f;
                            for (int i = 0; i < R; ++i) z += s_red[i];
                            named_bar_sync(1, nbar);
                        }
                        const bool bad = dead || !(z > 0.f) || !(z < 3.0e38f);
                        rz = bad ? 0.f : 1.0f / z;
                        if (tid == 0) {
                            if (bad) { s_flag[0] = 1; s_flag[1] = 1; }
                            if (!rev) p.nll[b] = bad ? 0.f : (float)(-((double)E0 + (double)log2f(z)) * kLn2);
                        }
                        if (wg) {
#pragma unroll
                            for (int q = 0; q < P; ++q) {
                                tB[q] *= rz;
                                ek[q][0] = tY[q] * rz;
                            }
                            store_row<P>(eb, tB);
                        }
                        end_step();
                        stp += row_step;
                        sto += row_step;
                        r0 = 1;
                    }
                    if (wg) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (r < r0) continue;
                            if (r >= rows) break;
                            float gB[P], gY[P];
#pragma unroll
                            for (int q = 0; q < P; ++q) { gB[q] = 0.f; gY[q] = 0.f; }
                            if (warp_on(tt0 + r)) {
                                float xs[P], ins[P];
                                fetch();
                                advance(yb_ptr, r * Vs, xs, ins);
                                if ((unsigned)u < win_cons) {
                                    // occupancy = a * p~ * 2^(off + o - E0) / z   (exact exponents); cells
                                    // outside [0, S] are exact zeros on my side, partner vectors are finite
                                    const float sb = pow2c(min(off + ob - E0, kLinHmax)) * rz;
                                    const float sy = hasX1 ? pow2c(min(off + oy - E0, kLinHmax)) * rz : 0.f;
#pragma unroll
                                    for (int q = 0; q < P; ++q) {
                                        gB[q] = aB[q] * (pb[q] * sb);
                                        if (q + 1 < P) gY[q] = aY[q] * (py[q] * sb);
                                    }
                                    if (hasX1) gY[P - 1] = aY[P - 1] * (py[P - 1] * sy);
                                }
                            }
#pragma unroll
                            for (int q = 0; q < P; ++q) ek[q][r * ER] = gY[q];
                            store_row<P>(eb + r * ER, gB);
                            stp += row_step;
                            sto += row_step;
                            end_step();
                            if (P < 4 && (r + 1) % P == 0 && r + 1 < rows) renorm();
                        }
                    }
                    renorm();
                    ring_part.advance();
                    e_buf ^= 1;
                }
                ring_y.advance();
            }
            __syncthreads();
            if (it == n1) {  // phase break (see the helper branch)
                cluster_sync_all();
                __syncthreads();
            }
        }
    } else {
        // =============================================================================
        // HELP: staging producer, fused softmax, gradient rows
        // =============================================================================
        const int half = lane >> 4, q16 = lane & 15;   // a helper handles two frames at a time
        int a_r0 = 0, a_c0 = lane;
        while (a_c0 >= V4) { a_c0 -= V4; ++a_r0; }
        const ptrdiff_t a_inc = (ptrdiff_t)tsign * (ptrdiff_t)frame_stride;
        auto issue_chunk = [&](int ka, int slot_a, int kp, int slot_p) {
            if (ka >= 0) {
                int tt0, rows;
                chunk_at(ka, tt0, rows);
                float* dst = s_y + (size_t)slot_a * TC * Vs;
                const float* src = acts_b + (ptrdiff_t)(tbase + tsign * tt0) * (ptrdiff_t)frame_stride;
                for (int r = a_r0, c = a_c0; r < rows;) {
                    cp_async16(dst + r * Vs + 4 * c, src + r * a_inc + 4 * c);
                    c += 32;
                    while (c >= V4) { c -= V4; ++r; }
                }
                cp_async_arrive(bar_acts + slot_a);
            }
            if (kp >= 0 && lane == 0) {
                int tt0, rows;
                chunk_at(kp, tt0, rows);
                uint64_t* bar = bar_part + slot_p;
                const int t_lo = rev ? tbase - (tt0 + rows - 1) : tt0;
                mbar_expect_tx(bar, (unsigned)(rows * RS) * 4u);
                bulk_g2s(s_stage + (size_t)slot_p * TC * RS, lat_b + (ptrdiff_t)t_lo * RS,
                         (unsigned)(rows * RS) * 4u, bar);
            }
        };

        // ---- fused softmax, in place, of two staged rows (one per half-warp) ------------
        auto softmax2 = [&](float* row, bool act) {
            float4* row4 = reinterpret_cast<float4*>(row);
            if (V4 <= 16) {  // the whole row is one float4 per lane of the half-warp
                float4 x = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
                if (q16 < V4) x = row4[q16];
                const float m = half_max(fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
                x.x = ex2f((x.x - m) * kLog2e); x.y = ex2f((x.y - m) * kLog2e);
                x.z = ex2f((x.z - m) * kLog2e); x.w = ex2f((x.w - m) * kLog2e);
                const float rs = 1.0f / half_sum((x.x + x.y) + (x.z + x.w));
                if (act && q16 < V4) row4[q16] = make_float4(x.x * rs, x.y * rs, x.z * rs, x.w * rs);
            } else {
                float m = -CUDART_INF_F, z = 0.f;
                for (int c = q16; c < V4; c += 16) {
                    const float4 x = row4[c];
                    m = fmaxf(m, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
                }
                m = half_max(m);
                for (int c = q16; c < V4; c += 16) {
                    float4 x = row4[c];
                    x.x = ex2f((x.x - m) * kLog2e); x.y = ex2f((x.y - m) * kLog2e);
                    x.z = ex2f((x.z - m) * kLog2e); x.w = ex2f((x.w - m) * kLog2e);
                    z += (x.x + x.y) + (x.z + x.w);
                    if (act) row4[c] = x;
                }
                const float rs = 1.0f / half_sum(z);
                __syncwarp();
                if (act)
                    for (int c = q16; c < V4; c += 16) {
                        const float4 x = row4[c];
                        row4[c] = make_float4(x.x * rs, x.y * rs, x.z * rs, x.w * rs);
                    }
            }
            if (act && q16 == 0) row[V] = 0.f;  // what padding pairs gather
        };

        // ---- gradient of two frames (one per half-warp) ---------------------------------
        //   e row = occupancies: [blank cells by slot: NP][label cells class-sorted, padded].
        //   The sum of class v is PS[cstart[v+1]] - PS[cstart[v]] of the exclusive prefix sums
        //   PS over the sorted label cells, which overwrite the occupancies in place.
        //   sum of ALL occupancies of a frame must be 1: the posterior-mass check.
        const int nb4 = (S + delta) / 4 + 1;                    // float4s covering blank slots 0..S+delta
        auto grad2 = [&](float* erow, const float* yrow, float* g, bool act) {
            const float4* eB4 = reinterpret_cast<const float4*>(erow);
            float bs = 0.f;
            for (int c = q16; c < nb4; c += 16) {
                const float4 x = eB4[c];
                bs += (x.x + x.y) + (x.z + x.w);
            }
            bs = half_sum(bs);
            float carry = 0.f;                                  // labels: 256 sorted cells per round
            float* eY = erow + NP;
            for (int base = 0; base < S; base += 256) {
                const int k0 = base + 16 * q16;                 // my 16 consecutive sorted cells
                float4* c4 = reinterpret_cast<float4*>(eY + ypad(k0));
                float o[16];
                const bool in = k0 < NP;                        // cells in [S, NP) hold 0
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (in) x = c4[j];
                    o[4 * j] = x.x; o[4 * j + 1] = x.y; o[4 * j + 2] = x.z; o[4 * j + 3] = x.w;
                }
                float run = 0.f;                                // exclusive prefix inside my 16 cells
#pragma unroll
                for (int j = 0; j < 16; ++j) { const float t = o[j]; o[j] = run; run += t; }
                float inc = run;                                // inclusive scan over the 16 lanes
#pragma unroll
                for (int s = 1; s < 16; s <<= 1) {
                    const float y = __shfl_up_sync(0xffffffffu, inc, s, 16);
                    if (q16 >= s) inc += y;
                }
                const float ex = carry + (inc - run);
                if (in && act) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        c4[j] = make_float4(ex + o[4 * j], ex + o[4 * j + 1], ex + o[4 * j + 2], ex + o[4 * j + 3]);
                }
                carry += __shfl_sync(0xffffffffu, inc, 15, 16);
            }
            __syncwarp();
            if (act) {
                if (!(fabsf((bs + carry) - 1.0f) <= kMassTol)) s_flag[0] = 1;   // NaN-safe
                for (int v = q16; v < V; v += 16) {
                    const int k0 = s_cstart[v], k1 = s_cstart[v + 1];
                    const float hi = (k1 < S) ? eY[ypad(k1)] : carry;   // PS[S] = total
                    const float lo = (k0 < S) ? eY[ypad(k0)] : carry;
                    const float occ = (hi - lo) + (v == blank ? bs : 0.f);
                    g[v] = gscale * (yrow[v] - occ);
                }
            }
        };

        // ---- the helper schedule (identical to ctc_pipe.cuh) -----------------------------
        int wgh_i = want_grad ? 1 : 0;
        asm volatile("" : "+r"(wgh_i));
        const bool wgh = wgh_i != 0;
        const bool iss_acts = H >= 3 ? hw == 0 : hw == H - 1, iss_part = hw == 0;
        const int n_sm = min(H, 2), n_gr = min(H, 2), gr_base = H - n_gr;
        const bool do_sm = hw < n_sm, do_gr = hw >= gr_base;
        const int sm_first = 2 * hw, sm_step = 2 * n_sm;
        const int gr_first = 2 * (hw - gr_base), gr_step = 2 * n_gr;
        Ring iss_a(NL), iss_p(NS), sm_a(NL), gr_a(NL);
        int gr_e = 0;
        if (iss_acts) {
            for (int k = 0; k <= D; ++k) {            // prologue: logits of chunks 0..D
                if (k < nch) issue_chunk(k, iss_a.slot, -1, 0);
                iss_a.advance();
            }
        }
        for (int it = 0; it < nch + 2; ++it) {
            {
                const int ka = it + D + 1, kp = it + D;
                const bool do_a = iss_acts && ka < nch;
                const bool do_p = iss_part && wgh && it >= n1 + 1 && kp < nch;
                if (do_a || do_p) issue_chunk(do_a ? ka : -1, iss_a.slot, do_p ? kp : -1, iss_p.slot);
                iss_a.advance();
                if (it >= n1 + 1) iss_p.advance();
            }
            const int kg = it - 2;
            if (kg >= 0) {
                if (do_gr && wgh && kg >= n1 && kg < nch && s_flag[1] == 0) {   // gradient rows of chunk it-2
                    int tt0, rows;
                    chunk_at(kg, tt0, rows);
                    for (int r0 = gr_first; r0 < rows; r0 += gr_step) {
                        const int r = min(r0 + half, rows - 1);
                        grad2(s_e + ((size_t)gr_e * TC + r) * ER, s_y + ((size_t)gr_a.slot * TC + r) * Vs,
                              grad_b + (size_t)(tbase + tsign * (tt0 + r)) * frame_stride, r0 + half < rows);
                    }
                }
                if (kg >= n1) gr_e ^= 1;
                gr_a.advance();
            }
            if (do_sm && it < nch) {                  // softmax of chunk `it`, two rows per pass
                int tt0, rows;
                chunk_at(it, tt0, rows);
                mbar_wait(bar_acts + sm_a.slot, sm_a.parity);
                float* base = s_y + (size_t)sm_a.slot * TC * Vs;
                for (int r0 = sm_first; r0 < rows; r0 += sm_step) {
                    const int r = r0 + half;
                    softmax2(base + min(r, rows - 1) * Vs, r < rows);
                }
            }
            sm_a.advance();
            __syncthreads();
            if (it == n1) {
                // Phase break: my REC warps have stored every row the partner will consume,
                // and (after the cluster barrier) vice versa.
                cluster_sync_all();
                if (iss_part) {
                    fence_proxy_async();
                    for (int k = n1; k <= n1 + D; ++k) {
                        if (k < nch && (wgh || k == n1)) issue_chunk(-1, 0, k, iss_p.slot);
                        iss_p.advance();
                    }
                }
                __syncthreads();
            }
        }
    }
    // every CTA reports whether its half passed; the log-domain kernel redoes flagged utterances
    __syncthreads();
    if (threadIdx.x == 0) flags[2 * b + (rev ? 1 : 0)] = s_flag[0];
}

}  // namespace ctcb200
