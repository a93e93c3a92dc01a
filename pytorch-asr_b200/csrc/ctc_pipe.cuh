// ctc_pipe.cuh -- the warp-specialised ("pipelined") fused CTC kernel for sm_100a.
//
// Same algorithm and HBM layout as ctc_fused_kernel (ctc_kernels.cuh): a 2-CTA
// cluster per utterance, the alpha CTA and the time/label-reversed beta CTA meet in
// the middle, ONE fp32 lattice goes through HBM (written once, read once), base-2
// log domain with exact per-warp integer offsets.  Here the CTA is split into roles
// so that the T-step dependent chain executes nothing but the lattice recursion:
//
//   warps [0, R)      REC   lattice recursion; P cell pairs per thread in registers;
//                           one named barrier per step among the R warps (none if R=1)
//   warps [R, R+H)    HELP  AHEAD of REC: TMA bulk copies (cp.async.bulk + mbarrier
//                           complete_tx) of the logit rows and of the partner's lattice
//                           rows into shared-memory rings; fused log_softmax in place.
//                           BEHIND REC: occupancies -> per-class sums (prefix sums over
//                           the class-sorted label cells) -> gradient rows; zero fill of
//                           rows t >= T_b.  A helper works on TWO frames at a time, one
//                           per half-warp, so the two dependent chains overlap.
//
// One CTA barrier per chunk of TC frames hands the rings over; consumers of TMA data
// wait on the ring slot's mbarrier.  Requires V % 4 == 0 (16-byte aligned logit rows).
//
// The kernel is latency-bound per warp (one in-order warp per role and SM
// sub-partition), so the code keeps every per-step quantity in a loop-carried register
// (pointers advance by a stride; nothing is re-derived from kernel parameters).
#pragma once
#ifndef CTC_LIN_PDL_EARLY
#define CTC_LIN_PDL_EARLY 1
#endif
#include "ctc_kernels.cuh"

namespace ctcb200 {

struct PipeParams {
    FusedParams f;
    int R, H;      // recursion / helper warps per CTA
    int NP;        // lattice pair slots = 32 * P * R
    int D;         // fetch distance in chunks
    int rotate;    // CTAs per launch "layer" (= #SMs): co-resident CTAs rotate their warp roles
    const int* redo;  // [2 * N] or nullptr: run only utterances the linear kernel flagged (ctc_lin.cuh)
    int utt_rot;      // linear kernel: cluster c works on utterance (c + utt_rot) mod n_utt (CTA placement)
    int map_mode;     // linear kernel: 0 = rotation only; 1 / 2 = length-balanced placement (lin_map_utt)
    int map_pairs;    // SM pairs of the device (clusters per launch "layer")
    int* queue;       // linear kernel, persistent launch: {next utterance, clusters that ran dry}; nullptr: one cluster per utterance
    int n_utt;        // utterances of this launch (queue mode)
};

// class-sorted label cell k lives at float index ypad(k) of the eY row: one float4 of
// padding per 16 cells makes "lane q owns cells [16q, 16q+16)" bank-conflict free
__host__ __device__ __forceinline__ int ypad(int k) { return k + ((k >> 4) << 2); }

struct PipeSmem {
    int lab, pos, cstart, fill, lp2, e, stage, bnd, red, ll, bars, total;  // byte offsets
    int Vs, ER, NL, NS;
    __host__ __device__ static int up(int x, int a) { return (x + a - 1) / a * a; }
    __host__ __device__ PipeSmem(int NP, int R, int V, int TC, int RS, int D, int ys = 0, int er = 0) {
        Vs = ys > 0 ? ys : up(V + 1, 4);   // ys: fixed row stride of the emission ring (ctc_lin.cuh)
        // [eB: NP][eY (class-sorted, padded)] + 4 floats of bank skew; er: occupancy row of ctc_lin.cuh
        ER = er > 0 ? er : NP + ypad(NP) + 4;
        NL = D + 4;              // lp2 ring: issued D+1 chunks early .. gradient 2 chunks later
        NS = D + 2;              // partner ring: issued D chunks early .. recursion 1 chunk later
        int o = 0;
        lab = o;    o += up(NP * 4, 16);
        pos = o;    o += up(NP * 4, 16);
        cstart = o; o += up((V + 2) * 4, 16);
        fill = o;   o += up((V + 2) * 4, 16);
        lp2 = o;    o += up(NL * TC * Vs * 4, 16);
        e = o;      o += up(2 * TC * ER * 4, 16);
        stage = o;  o += up(NS * TC * RS * 4, 16);
        bnd = o;    o += up(2 * (R + 1) * 8, 16);
        red = o;    o += up(32 * 4, 16);
        ll = o;     o += 16;
        bars = o;   o += up((NL + NS) * 8, 16);
        total = o;
    }
};

__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned ok = 0;
    for (unsigned spins = 0; !ok; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (spins > (1u << 24)) __trap();  // a lost transaction must fail loudly, not hang
    }
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (UBLKCP in SASS)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// make `bar` observe the completion of all cp.async copies this thread issued so far; counts
// as one of the barrier's expected arrivals (the logits barriers expect 32: one per lane)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// sum / max over the 16 lanes of a half-warp (xor offsets < 16 stay inside the half)
__device__ __forceinline__ float half_sum(float x) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ float half_max(float x) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

struct Ring {  // incrementally maintained (slot, phase parity) of a ring of n mbarrier slots
    int slot, n;
    unsigned parity;
    __device__ __forceinline__ Ring(int n_) : slot(0), n(n_), parity(0) {}
    __device__ __forceinline__ void advance() { if (++slot == n) { slot = 0; parity ^= 1u; } }
};

template <int P, int MAXT, int MINB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MAXT, MINB)
ctc_pipe_kernel(const PipeParams pp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FusedParams& p = pp.f;
    const int NT = blockDim.x, NW = NT >> 5;
    const int R = pp.R, H = pp.H, NP = pp.NP;
    // Warps map to SM sub-partitions by (physical warp id % 4).  CTAs that end up on the
    // same SM (launch "layers" of `rotate` CTAs) rotate their roles so that the MUFU-heavy
    // recursion warps do not all land on the same sub-partitions.
    const int lane = threadIdx.x & 31;
    const int shift = pp.rotate > 0 ? (int)((blockIdx.x / pp.rotate) * R) % NW : 0;
    const int w = ((int)(threadIdx.x >> 5) + NW - shift) % NW;   // role (virtual) warp id
    const int tid = w * 32 + lane;
    const int b = p.utt_begin + (blockIdx.x >> 1);
    const bool rev = (blockIdx.x & 1) != 0;
    // fallback mode: both CTAs of the cluster leave unless the linear kernel flagged the utterance
    // (launched with programmatic stream serialization: wait for the linear kernel's flags)
#if CTC_LIN_PDL_EARLY
    if (pp.redo != nullptr) asm volatile("griddepcontrol.launch_dependents;");   // (the loss reduction may queue up behind)
#endif
    if (pp.redo != nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (pp.redo != nullptr && (pp.redo[2 * b] | pp.redo[2 * b + 1]) == 0) return;
    const int T = p.T, V = p.V, blank = p.blank;
    const int RS = p.row_stride, TC = p.chunk, D = pp.D;
    const Clamp clp{p.use_clamp != 0, p.clamp_lo, p.clamp_hi};
    const bool is_rec = w < R;
    const int hw = w - R;  // helper index (>= 0 for helpers)

    const PipeSmem lay(NP, R, V, TC, RS, D);
    const int Vs = lay.Vs, ER = lay.ER, NL = lay.NL, NS = lay.NS;
    int* s_lab = reinterpret_cast<int*>(smem_raw + lay.lab);
    int* s_pos = reinterpret_cast<int*>(smem_raw + lay.pos);
    int* s_cstart = reinterpret_cast<int*>(smem_raw + lay.cstart);
    int* s_fill = reinterpret_cast<int*>(smem_raw + lay.fill);
    float* s_lp2 = reinterpret_cast<float*>(smem_raw + lay.lp2);
    float* s_e = reinterpret_cast<float*>(smem_raw + lay.e);
    float* s_stage = reinterpret_cast<float*>(smem_raw + lay.stage);
    float2* s_bnd = reinterpret_cast<float2*>(smem_raw + lay.bnd);
    float* s_red = reinterpret_cast<float*>(smem_raw + lay.red);
    float* s_ll = reinterpret_cast<float*>(smem_raw + lay.ll);  // [2] = infeasible flag
    uint64_t* bar_acts = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);   // [NL]
    uint64_t* bar_part = bar_acts + NL;                                      // [NS]

    int Tb = p.in_lens[b], S = p.tgt_lens[b];
    if (Tb < 0 || Tb > T || S < 0 || S > NP - 1) {
        if (tid == 0) atomicOr(p.status, kStatusBadLength);
        Tb = min(max(Tb, 0), T);
        S = min(max(S, 0), NP - 1);
    }
    const int32_t* tg = p.targets + p.tgt_off[b];
    const bool want_grad = p.grad != nullptr;
    const float gscale = p.grad_scale ? p.grad_scale[b] : 1.0f;
    const size_t frame_stride = (size_t)p.frame_stride;
    const float* acts_b = p.acts + (size_t)b * (size_t)p.utt_stride;
    float* grad_b = want_grad ? p.grad + (size_t)b * (size_t)p.utt_stride : nullptr;
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
        if (!rev && tid == 0)
            p.nll[b] = (S == 0 || p.zero_infinity) ? 0.0f : CUDART_INF_F;
        return;  // both CTAs of the cluster take this exit
    }

    // ---- per-utterance setup (all warps) --------------------------------------------
    for (int i = tid; i < NP; i += NT) {
        int c = V;  // padding pairs gather the kNeg slot of the lp2 row
        if (i < S) {
            c = rev ? tg[S - 1 - i] : tg[i];
            if (c < 0 || c >= V) {
                atomicOr(p.status, kStatusBadLabel);
                c = min(max(c, 0), V - 1);
            }
        }
        s_lab[i] = c;
        s_pos[i] = i;  // padding pairs park their kNeg at sorted slot i >= S
    }
    for (int v = tid; v < V + 2; v += NT) s_cstart[v] = 0;
    for (int i = tid; i < 2 * (R + 1); i += NT) s_bnd[i] = make_float2(kNeg, 0.f);
    if (tid == 0) {
        s_ll[0] = 0.f; s_ll[1] = 0.f; s_ll[2] = 0.f;
        for (int i = 0; i < NL; ++i) mbar_init(bar_acts + i, 32);   // 32 lanes' cp.async
        for (int i = 0; i < NS; ++i) mbar_init(bar_part + i, 1);    // one TMA producer
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();
    if (want_grad)
        for (int i = tid; i < S; i += NT) atomicAdd(&s_cstart[s_lab[i] + 1], 1);
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
        // class-sorted position of every label, equal labels in sweep order (deterministic):
        // 32 labels per round, lanes with the same class rank themselves with match.any
        for (int base = 0; base < S; base += 32) {
            const int i = base + lane;
            const int c = (i < S) ? s_lab[i] : -1 - lane;
            const unsigned peers = __match_any_sync(0xffffffffu, c);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            int first = 0;
            if (i < S) first = s_fill[c];
            __syncwarp();
            if (i < S) {
                s_pos[i] = first + rank;
                if (rank == 0) s_fill[c] = first + __popc(peers);
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- sweep geometry ---------------------------------------------------------------
    // Pair i is inside the band at sweep step tt iff 0 <= tt - i <= C (C = T_b - S): it
    // can be reached from the start and can still finish.  Outside the band a cell is
    // kNeg on my side or kNeg on the partner's side, so combining needs no per-cell test
    // as long as whatever is READ has been WRITTEN: every store / staging / activity
    // window below is widened by P pairs, so a consumer thread with one cell in the band
    // finds all the partner vectors its 2P cells map to.
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

#ifdef CTC_B200_PROFILE
    // developer instrumentation: busy cycles of CTA 0 per role -> workspace header (u64 at +64):
    // [0] REC warp 0, [1] helper 0, [2] helper 1, [3] wall, [4] iterations, [5] T_b
    unsigned long long* prof = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(p.status) + 64);
    const bool prof_on = blockIdx.x == 0 && lane == 0 && (w == 0 || w == R || w == R + 1);
    const int prof_slot = w == 0 ? 0 : (w == R ? 1 : 2);
    long long prof_t0 = clock64();
    const long long prof_start = prof_t0;
#define PROF_BEGIN() do { prof_t0 = clock64(); } while (0)
#define PROF_END() do { if (prof_on) atomicAdd(prof + prof_slot, (unsigned long long)(clock64() - prof_t0)); } while (0)
#define PROF_SEC(slot) do { if (prof_on && w == R) atomicAdd(prof + (slot), (unsigned long long)(clock64() - prof_t1)); prof_t1 = clock64(); } while (0)
    long long prof_t1 = prof_t0;
#else
#define PROF_SEC(slot) do {} while (0)
#define PROF_BEGIN() do {} while (0)
#define PROF_END() do {} while (0)
#endif

    if (is_rec) {
        // =============================================================================
        // REC
        // =============================================================================
        const int i0 = tid * P;
        int lab[P], ysl[P];
        bool skip[P], selB[P], selY[P], vB[P], vY[P];
        float aB[P], aY[P];
        const int jB0 = S - i0;                            // partner pair index of my first blank
        const int pw_hi = max(jB0, 0) / (32 * P);          // at most two partner warps per thread
        const int pw_lo = max(pw_hi - 1, 0);
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const int i = i0 + k;
            lab[k] = s_lab[i];
            skip[k] = (i >= 1 && i < S && lab[k] != s_lab[i - 1]);
            ysl[k] = NP + ypad(s_pos[i]);                  // where my label cell goes in an e row
            selB[k] = max(jB0 - k, 0) / (32 * P) == pw_hi;
            selY[k] = max(jB0 - 1 - k, 0) / (32 * P) == pw_hi;
            vB[k] = i <= S;
            vY[k] = i < S;
            aB[k] = kNeg;
            aY[k] = kNeg;
        }
        if (tid == 0) aB[0] = 0.0f;  // virtual row "-1": log(1) in front of the first blank
        float off = 0.0f;            // this warp's exact integer offset (true = off + a)
        bool fresh = (w != 0);       // warp has not received any real value yet
        float ll_int = 0.f, ll_frac = 0.f;
        // thread / warp activity windows in sweep steps (widened by P pairs, see above)
        const unsigned win_store = (i0 - P <= S) ? (unsigned)(C + 3 * P) : 0u;   // (tt - i0 + P) < win
        const unsigned win_cons = (i0 <= S) ? (unsigned)(C + P) : 0u;            // (tt - i0) < win
        const int w_first = (32 * P * w - P <= S) ? 32 * P * w - P : 0x3fffffff;
        const int w_last = C + 32 * P * w + 32 * P - 1 + P;
        // loop-carried shared-memory pointers for the warp-to-warp boundary (double buffered)
        float2* bnd_rd = s_bnd + w;              // slot 0 is the constant (kNeg, 0)
        float2* bnd_wr = s_bnd + (R + 1) + w + 1;
        const bool lane0 = lane == 0, lane31 = (lane == 31) && R > 1;
        const int nbar = 32 * R;
        const ptrdiff_t row_inc = (ptrdiff_t)tsign * RS;

        // keep per-step constants in registers (the compiler would otherwise re-derive them
        // from the kernel parameters inside the loops; every such reload is an exposed
        // latency for a single in-order warp)
        int wg_i = want_grad ? 1 : 0, vs_i = Vs, rs_i = RS, er_i = ER, np_i = NP;
        asm volatile("" : "+r"(wg_i), "+r"(vs_i), "+r"(rs_i), "+r"(er_i), "+r"(np_i));
        const bool wg = wg_i != 0;
        const ptrdiff_t row_step = (ptrdiff_t)tsign * rs_i;

        auto renorm = [&]() {
            // exact integer renormalisation keeps |a| small (fp32 accuracy at long T)
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
        };
        // one recursion step on registers: a[t] <- a[t-1].  STEADY: the warp has real values
        // and renormalises at chunk ends, so neither test is compiled into the step.
        auto advance = [&](auto steady_tag, int tt, const float* lp2, float& lpb, float (&lpl)[P]) {
            constexpr bool STEADY = decltype(steady_tag)::value;
            lpb = lp2[blank];
            float am1 = __shfl_up_sync(0xffffffffu, aY[P - 1], 1);
            float2 bq = make_float2(kNeg, 0.f);
            if (lane0) bq = *bnd_rd;
            if (!STEADY && fresh) {
                const float bv = __shfl_sync(0xffffffffu, bq.x, 0);
                const float bo = __shfl_sync(0xffffffffu, bq.y, 0);
                if (bv > kRealThresh) { off = bo; fresh = false; }
            }
            if (lane0) am1 = bq.x + (bq.y - off);
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
            if (!STEADY && (tt & kRenormMask) == kRenormMask) renorm();
            if (lane31) *bnd_wr = make_float2(aY[P - 1], off);
        };
        auto end_step = [&]() {
            float2* t = bnd_rd; bnd_rd = bnd_wr - 1; bnd_wr = t + 1;   // flip the double buffer
            if (R > 1) named_bar_sync(1, nbar);
        };
        using TrueT = std::integral_constant<bool, true>;
        using FalseT = std::integral_constant<bool, false>;

        Ring ring_lp2(NL), ring_part(NS);   // position of the chunk REC works on
        int e_buf = 0;
        for (int it = 0; it < nch + 2; ++it) {
            PROF_BEGIN();
            const int k = it - 1;
            if (k >= 0 && k < nch) {
                int tt0, rows;
                chunk_at(k, tt0, rows);
                const int tt_end = tt0 + rows - 1;
                // steady chunk: the warp is active for every row and already holds real values
                const bool steady = wg && !fresh && tt0 >= w_first && tt_end <= w_last;
                const bool renorm_after = ((tt_end + 1) >> 3) != (tt0 >> 3);
                const float* lp2 = s_lp2 + (size_t)ring_lp2.slot * TC * vs_i;
                if (k < n1) {
                    // ---- store chunk: rows go to HBM for the partner --------------------
                    float* row = lat_b + (ptrdiff_t)(tbase + tsign * tt0) * rs_i + i0;
                    float* offp = lat_b + (ptrdiff_t)(tbase + tsign * tt0) * rs_i + 2 * np_i + w;
                    if (steady) {
                        // whole rows of the warp are stored (cells outside the band hold kNeg
                        // or finite junk nobody reads): no per-thread test on the step
                        for (int r = 0; r < rows; ++r, lp2 += vs_i, row += row_step, offp += row_step) {
                            float lpb, lpl[P];
                            advance(TrueT{}, tt0 + r, lp2, lpb, lpl);
                            store_vec<P>(row, aB);
                            store_vec<P>(row + np_i, aY);
                            if (lane0) *offp = off;
                            end_step();
                        }
                        if (renorm_after) renorm();
                    } else {
                        for (int r = 0; r < rows; ++r, lp2 += vs_i, row += row_step, offp += row_step) {
                            const int tt = tt0 + r;
                            if (tt >= w_first && tt <= w_last) {
                                float lpb, lpl[P];
                                advance(FalseT{}, tt, lp2, lpb, lpl);
                                if (wg || tt == n_store - 1) {
                                    if ((unsigned)(tt - i0 + P) < win_store) {
                                        store_vec<P>(row, aB);
                                        store_vec<P>(row + np_i, aY);
                                    }
                                    if (lane0) *offp = off;
                                }
                            }
                            end_step();
                        }
                    }
                } else {
                    // ---- consume chunk: combine with the partner's stored rows ----------
                    mbar_wait(bar_part + ring_part.slot, ring_part.parity);   // TMA data landed
                    // the chunk was staged as one block in frame order: for the reversed sweep
                    // its rows run backwards, so start at the last one and step by -RS
                    const float* st = s_stage + ((size_t)ring_part.slot * TC + (rev ? rows - 1 : 0)) * rs_i;
                    const int st_step = (int)row_step;
                    float* erow = s_e + (size_t)e_buf * TC * er_i;
                    const float* stp = st + jB0;                 // partner blank of my pair q: stp[-q]
                    const float* sto_hi = st + 2 * np_i + pw_hi; // partner warp offsets
                    const float* sto_lo = st + 2 * np_i + pw_lo;
                    int r = 0;
                    if (k == n1) {
                        // first combined row: also yields the log-likelihood.  ll = ll_int +
                        // ll_frac with an exact integer pivot: ll_int = max(n + rint(e)),
                        // ll_frac = log2 sum 2^(e + n - ll_int)
                        const int tt = tt0;
                        float eBv[P], eYv[P], nBv[P], nYv[P];
#pragma unroll
                        for (int q = 0; q < P; ++q) { eBv[q] = kNeg; eYv[q] = kNeg; nBv[q] = 0.f; nYv[q] = 0.f; }
                        if (tt >= w_first && tt <= w_last) {
                            float lpb, lpl[P];
                            advance(FalseT{}, tt, lp2, lpb, lpl);
                            if ((unsigned)(tt - i0) < win_cons) {
                                const float nhi = off + *sto_hi, nlo = off + *sto_lo;
#pragma unroll
                                for (int q = 0; q < P; ++q) {
                                    if (vB[q]) { eBv[q] = (aB[q] + stp[-q]) - lpb; nBv[q] = selB[q] ? nhi : nlo; }
                                    if (vY[q]) { eYv[q] = (aY[q] + stp[np_i - 1 - q]) - lpl[q]; nYv[q] = selY[q] ? nhi : nlo; }
                                }
                            }
                        }
                        float pm = kNeg;
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            if (eBv[q] > kRealThresh) pm = fmaxf(pm, nBv[q] + rintf(eBv[q]));
                            if (eYv[q] > kRealThresh) pm = fmaxf(pm, nYv[q] + rintf(eYv[q]));
                        }
                        pm = warp_max(pm);
                        if (R > 1) {
                            if (lane0) s_red[w] = pm;
                            named_bar_sync(1, nbar);
                            for (int i = 0; i < R; ++i) pm = fmaxf(pm, s_red[i]);
                            named_bar_sync(1, nbar);
                        }
                        const bool infeasible = !(pm > kRealThresh);
                        ll_int = infeasible ? 0.f : pm;
                        float z = 0.f;
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            eBv[q] = (eBv[q] > kRealThresh) ? eBv[q] + (nBv[q] - ll_int) : kNeg;
                            eYv[q] = (eYv[q] > kRealThresh) ? eYv[q] + (nYv[q] - ll_int) : kNeg;
                            z += ex2f(eBv[q]) + ex2f(eYv[q]);
                        }
                        z = warp_sum(z);
                        if (R > 1) {
                            if (lane0) s_red[w] = z;
                            named_bar_sync(1, nbar);
                            z = 0.f;
                            for (int i = 0; i < R; ++i) z += s_red[i];
                            named_bar_sync(1, nbar);
                        }
                        ll_frac = infeasible ? 0.f : lg2f(z);
                        if (tid == 0) {
                            s_ll[2] = infeasible ? 1.f : 0.f;
                            if (!rev) {
                                float out;
                                if (infeasible) out = p.zero_infinity ? 0.0f : CUDART_INF_F;
                                else out = (float)(-((double)ll_int + (double)ll_frac) * kLn2);
                                p.nll[b] = out;
                            }
                        }
                        if (wg) {
#pragma unroll
                            for (int q = 0; q < P; ++q) {
                                eBv[q] -= ll_frac;
                                erow[ysl[q]] = eYv[q] - ll_frac;
                            }
                            store_vec<P>(erow + i0, eBv);
                        }
                        end_step();
                        r = 1; lp2 += vs_i; stp += st_step; sto_hi += st_step; sto_lo += st_step; erow += er_i;
                    }
                    if (wg) {
                        // a + partner - lp + (offsets - ll): the integer parts combine exactly
                        // The partner values are loaded BEFORE the recursion step (they do not depend
                        // on it), so their shared-memory latency hides behind the MUFU chain.
                        float pb[P], py[P], ohi, olo;
                        auto fetch = [&]() {
#pragma unroll
                            for (int q = 0; q < P; ++q) { pb[q] = stp[-q]; py[q] = stp[np_i - 1 - q]; }
                            ohi = *sto_hi;
                            olo = *sto_lo;
                        };
                        auto combine = [&](int tt, float lpb, const float (&lpl)[P], float (&eBv)[P], float (&eYv)[P]) {
                            if ((unsigned)(tt - i0) < win_cons) {
                                const float chi = ((off + ohi) - ll_int) - ll_frac;
                                const float clo = ((off + olo) - ll_int) - ll_frac;
                                const float bhi = chi - lpb, blo = clo - lpb;
#pragma unroll
                                for (int q = 0; q < P; ++q) {   // padding pairs may have loaded junk: vB / vY guard them
                                    if (vB[q]) eBv[q] = (aB[q] + pb[q]) + (selB[q] ? bhi : blo);
                                    if (vY[q]) eYv[q] = (aY[q] + py[q]) + ((selY[q] ? chi : clo) - lpl[q]);
                                }
                            }
                        };
                        if (steady) {
                            for (; r < rows; ++r, lp2 += vs_i, stp += st_step, sto_hi += st_step, sto_lo += st_step, erow += er_i) {
                                float eBv[P], eYv[P], lpb, lpl[P];
#pragma unroll
                                for (int q = 0; q < P; ++q) { eBv[q] = kNeg; eYv[q] = kNeg; }
                                fetch();
                                advance(TrueT{}, tt0 + r, lp2, lpb, lpl);
                                combine(tt0 + r, lpb, lpl, eBv, eYv);
#pragma unroll
                                for (int q = 0; q < P; ++q) erow[ysl[q]] = eYv[q];
                                store_vec<P>(erow + i0, eBv);
                                end_step();
                            }
                            if (renorm_after) renorm();
                        } else {
                            for (; r < rows; ++r, lp2 += vs_i, stp += st_step, sto_hi += st_step, sto_lo += st_step, erow += er_i) {
                                const int tt = tt0 + r;
                                float eBv[P], eYv[P];
#pragma unroll
                                for (int q = 0; q < P; ++q) { eBv[q] = kNeg; eYv[q] = kNeg; }
                                if (tt >= w_first && tt <= w_last) {
                                    float lpb, lpl[P];
                                    fetch();
                                    advance(FalseT{}, tt, lp2, lpb, lpl);
                                    combine(tt, lpb, lpl, eBv, eYv);
                                }
#pragma unroll
                                for (int q = 0; q < P; ++q) erow[ysl[q]] = eYv[q];
                                store_vec<P>(erow + i0, eBv);
                                end_step();
                            }
                        }
                    }
                    ring_part.advance();
                    e_buf ^= 1;
                }
                ring_lp2.advance();
            }
            PROF_END();
            __syncthreads();
            if (it == n1) {  // phase break (see the helper branch)
                cluster_sync_all();
                __syncthreads();
            }
        }
    } else {
        // =============================================================================
        // HELP: TMA producer, fused log_softmax, gradient rows
        // =============================================================================
        const int half = lane >> 4, q16 = lane & 15;   // a helper handles two frames at a time

        // ---- staging of one chunk -------------------------------------------------------------
        //   logits rows of chunk ka : cp.async (LDGSTS, 16 B per lane and copy) by all 32 lanes,
        //                             completion signalled on the slot's mbarrier
        //   partner rows of chunk kp: ONE TMA bulk copy (the rows of a chunk are contiguous in
        //                             the lattice); regions the partner never wrote are copied
        //                             but never read
        // per-lane (row, float4) position of its first logits copy; later copies are 32 float4 on
        int a_r0 = 0, a_c0 = lane;
        while (a_c0 >= V4) { a_c0 -= V4; ++a_r0; }
        const ptrdiff_t a_inc = (ptrdiff_t)tsign * (ptrdiff_t)frame_stride;
        auto issue_chunk = [&](int ka, int slot_a, int kp, int slot_p) {
            if (ka >= 0) {
                int tt0, rows;
                chunk_at(ka, tt0, rows);
                float* dst = s_lp2 + (size_t)slot_a * TC * Vs;
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

        // ---- fused log_softmax, in place, of two staged rows (one per half-warp) -------
        auto softmax2 = [&](float* row, bool act) {
            float4* row4 = reinterpret_cast<float4*>(row);
            if (V4 <= 16) {  // the whole row is one float4 per lane of the half-warp
                float4 x = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
                if (q16 < V4) x = row4[q16];
                const float4 raw = x;
                if (q16 < V4) { x.x = clp.cin(x.x); x.y = clp.cin(x.y); x.z = clp.cin(x.z); x.w = clp.cin(x.w); }
                const float m = half_max(fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
                x.x = (x.x - m) * kLog2e; x.y = (x.y - m) * kLog2e;
                x.z = (x.z - m) * kLog2e; x.w = (x.w - m) * kLog2e;
                const float lz = lg2f(half_sum((ex2f(x.x) + ex2f(x.y)) + (ex2f(x.z) + ex2f(x.w))));
                if (act && q16 < V4)
                    row4[q16] = make_float4(clp.tag(fmaxf(x.x - lz, kNeg), raw.x), clp.tag(fmaxf(x.y - lz, kNeg), raw.y),
                                            clp.tag(fmaxf(x.z - lz, kNeg), raw.z), clp.tag(fmaxf(x.w - lz, kNeg), raw.w));
            } else {
                float m = -CUDART_INF_F, z = 0.f;
                for (int c = q16; c < V4; c += 16) {
                    const float4 x = row4[c];
                    m = fmaxf(m, fmaxf(fmaxf(clp.cin(x.x), clp.cin(x.y)), fmaxf(clp.cin(x.z), clp.cin(x.w))));
                }
                m = half_max(m);
                for (int c = q16; c < V4; c += 16) {
                    const float4 x = row4[c];
                    z += (ex2f((clp.cin(x.x) - m) * kLog2e) + ex2f((clp.cin(x.y) - m) * kLog2e)) +
                         (ex2f((clp.cin(x.z) - m) * kLog2e) + ex2f((clp.cin(x.w) - m) * kLog2e));
                }
                const float lz = lg2f(half_sum(z));
                if (act)
                    for (int c = q16; c < V4; c += 16) {
                        const float4 x = row4[c];
                        row4[c] = make_float4(clp.tag(fmaxf((clp.cin(x.x) - m) * kLog2e - lz, kNeg), x.x),
                                              clp.tag(fmaxf((clp.cin(x.y) - m) * kLog2e - lz, kNeg), x.y),
                                              clp.tag(fmaxf((clp.cin(x.z) - m) * kLog2e - lz, kNeg), x.z),
                                              clp.tag(fmaxf((clp.cin(x.w) - m) * kLog2e - lz, kNeg), x.w));
                    }
            }
            if (act && q16 == 0) row[V] = kNeg;  // what padding pairs gather
        };

        // ---- gradient of two frames (one per half-warp) ---------------------------------
        //   blank cells: plain sum.  Label cells are stored class-sorted (padded layout
        //   ypad), lane q of the half owns 16 consecutive sorted cells per 256-cell round;
        //   the sum of class v is PS[cstart[v+1]] - PS[cstart[v]] of the exclusive prefix
        //   sums PS, which overwrite the occupancy exponents in place.
        auto grad2 = [&](float* erow, const float* lp2row, float* g, bool act, bool infeasible) {
            if (infeasible) {
                const float fill = p.zero_infinity ? 0.0f : CUDART_NAN_F;
                if (act) for (int v = q16; v < V; v += 16) g[v] = fill;
                return;
            }
            const float4* eB4 = reinterpret_cast<const float4*>(erow);
            float bs = 0.f;
            for (int c = q16; c * 4 <= S; c += 16) {            // blanks 0..S (padding holds kNeg)
                const float4 x = eB4[c];
                bs += (ex2f(x.x) + ex2f(x.y)) + (ex2f(x.z) + ex2f(x.w));
            }
            bs = half_sum(bs);
            float carry = 0.f;                                  // labels: 256 sorted cells per round
            float* eY = erow + NP;
            for (int base = 0; base < S; base += 256) {
                const int k0 = base + 16 * q16;                 // my 16 consecutive sorted cells
                float4* c4 = reinterpret_cast<float4*>(eY + ypad(k0));
                float o[16];
                const bool in = k0 < NP;                        // cells in [S, NP) hold kNeg -> 0
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 x = make_float4(kNeg, kNeg, kNeg, kNeg);
                    if (in) x = c4[j];
                    o[4 * j] = ex2f(x.x); o[4 * j + 1] = ex2f(x.y); o[4 * j + 2] = ex2f(x.z); o[4 * j + 3] = ex2f(x.w);
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
            if (act)
                for (int v = q16; v < V; v += 16) {
                    const int k0 = s_cstart[v], k1 = s_cstart[v + 1];
                    const float hi = (k1 < S) ? eY[ypad(k1)] : carry;   // PS[S] = total
                    const float lo = (k0 < S) ? eY[ypad(k0)] : carry;
                    const float occ = (hi - lo) + (v == blank ? bs : 0.f);
                    const float lpv = lp2row[v];
                    g[v] = clp.tagged(lpv) ? 0.0f : gscale * (ex2f(lpv) - occ);
                }
        };

        // ---- the helper schedule ---------------------------------------------------------
        // Iteration `it`:  issue TMA { logits of chunk it+D+1, partner rows of chunk it+D };
        //                  softmax chunk it;  gradient rows of chunk it-2.
        // (REC runs chunk it-1.)  Partner rows of the first D+1 consume chunks cannot be
        // requested before the partner CTA wrote them: they are issued at the phase break.
        // The last helper issues the logit copies, helper 0 the partner copies.
        // Work of one iteration = TMA issue + softmax passes + gradient passes (a pass = two
        // rows).  Dealing: H=1 all; H=2 (TMA partner, S0, G0 | TMA logits, S1, G1);
        // H>=3 (TMA, S0 | G0, S1 | G1 ...): softmax passes go to helpers 0/1, gradient passes
        // to the last two helpers, the TMA issue to helper 0.
        int wgh_i = want_grad ? 1 : 0;
        asm volatile("" : "+r"(wgh_i));                  // keep the flag in a register
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
            PROF_BEGIN();
            PROF_SEC(13);
            {
                const int ka = it + D + 1, kp = it + D;
                const bool do_a = iss_acts && ka < nch;
                const bool do_p = iss_part && wgh && it >= n1 + 1 && kp < nch;
                if (do_a || do_p) issue_chunk(do_a ? ka : -1, iss_a.slot, do_p ? kp : -1, iss_p.slot);
                iss_a.advance();
                if (it >= n1 + 1) iss_p.advance();
            }
            // gradient first: its inputs are on chip already, the logits TMA gets more time
            const int kg = it - 2;
            if (kg >= 0) {
                if (do_gr && wgh && kg >= n1 && kg < nch) {   // gradient rows of chunk it-2
                    int tt0, rows;
                    chunk_at(kg, tt0, rows);
                    const bool infeasible = s_ll[2] != 0.f;
                    for (int r0 = gr_first; r0 < rows; r0 += gr_step) {
                        const int r = min(r0 + half, rows - 1);
                        grad2(s_e + ((size_t)gr_e * TC + r) * ER, s_lp2 + ((size_t)gr_a.slot * TC + r) * Vs,
                              grad_b + (size_t)(tbase + tsign * (tt0 + r)) * frame_stride,
                              r0 + half < rows, infeasible);
                    }
                    gr_e ^= 1;
                }
                gr_a.advance();
            }
            PROF_SEC(10);
            if (do_sm && it < nch) {                  // softmax of chunk `it`, two rows per pass
                int tt0, rows;
                chunk_at(it, tt0, rows);
                mbar_wait(bar_acts + sm_a.slot, sm_a.parity);
                PROF_SEC(7);
                float* base = s_lp2 + (size_t)sm_a.slot * TC * Vs;
                for (int r0 = sm_first; r0 < rows; r0 += sm_step) {
                    const int r = r0 + half;
                    softmax2(base + min(r, rows - 1) * Vs, r < rows);
                }
                PROF_SEC(8);
            }
            sm_a.advance();
            PROF_END();
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
#ifdef CTC_B200_PROFILE
    if (blockIdx.x == 0 && tid == 0) {
        prof[3] = (unsigned long long)(clock64() - prof_start);
        prof[4] = (unsigned long long)(nch + 2);
        prof[5] = (unsigned long long)Tb;
    }
#endif
}

}  // namespace ctcb200
