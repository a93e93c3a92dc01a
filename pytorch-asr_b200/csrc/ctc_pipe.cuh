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
//   warps [R, R+H)    HELP  one chunk AHEAD: TMA bulk copies (cp.async.bulk +
//                           mbarrier complete_tx) of the logit rows and of the
//                           partner's lattice rows into shared-memory rings, fused
//                           log_softmax in place;
//                           two chunks BEHIND: occupancies -> per-class sums (prefix
//                           sums over the class-sorted label cells) -> gradient rows,
//                           written with coalesced 128-bit stores; zero fill of rows
//                           t >= T_b
//
// One CTA barrier per chunk of TC frames hands the rings over; consumers of TMA data
// wait on the ring slot's mbarrier.  Requires V % 4 == 0 (16-byte aligned logit rows).
#pragma once
#include "ctc_kernels.cuh"

namespace ctcb200 {

struct PipeParams {
    FusedParams f;
    int R, H;      // recursion / helper warps per CTA
    int NP;        // lattice pair slots = 32 * P * R
    int D;         // fetch distance in chunks
    int rotate;    // CTAs per launch "layer" (= #SMs): co-resident CTAs rotate their warp roles
};

struct PipeSmem {
    int lab, cstart, lp2, e, stage, bnd, red, ll, bars, total;  // byte offsets
    int Vs, ER, NL, NS;
    __host__ __device__ static int up(int x, int a) { return (x + a - 1) / a * a; }
    __host__ __device__ PipeSmem(int NP, int R, int V, int TC, int RS, int D) {
        Vs = up(V + 1, 4);
        ER = 2 * NP + 4;  // [eB: NP][eY (class-sorted): NP] + 4 floats of bank skew per row
        NL = D + 4;       // lp2 ring: issued D+1 chunks early .. gradient 2 chunks later
        NS = D + 2;       // partner ring: issued D chunks early .. recursion 1 chunk later
        int o = 0;
        lab = o;    o += up(NP * 4, 16);
        cstart = o; o += up((V + 2) * 4, 16);
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

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
    const int T = p.T, N = p.N, V = p.V, blank = p.blank;
    const int RS = p.row_stride, TC = p.chunk, D = pp.D;
    const bool is_rec = w < R;
    const int hw = w - R;  // helper index (>= 0 for helpers)

    const PipeSmem lay(NP, R, V, TC, RS, D);
    const int Vs = lay.Vs, ER = lay.ER, NL = lay.NL, NS = lay.NS;
    int* s_lab = reinterpret_cast<int*>(smem_raw + lay.lab);
    int* s_cstart = reinterpret_cast<int*>(smem_raw + lay.cstart);
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
    const size_t frame_stride = (size_t)N * V;
    const float* acts_b = p.acts + (size_t)b * V;
    float* grad_b = want_grad ? p.grad + (size_t)b * V : nullptr;
    const int V4 = V >> 2;

    // ---- helpers: mandatory zero fill of gradient rows t >= T_b (no compute) --------
    if (want_grad && !is_rec) {
        const int nrows = T - Tb;
        const int mine = (nrows + (rev ? 0 : 1)) >> 1;  // rows Tb+rev, Tb+rev+2, ...
        const int first = Tb + (rev ? 1 : 0);
        for (int r = hw; r < mine; r += H) {
            float4* g4 = reinterpret_cast<float4*>(grad_b + (size_t)(first + 2 * r) * frame_stride);
            for (int c = lane; c < V4; c += 32) g4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
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
    }
    for (int v = tid; v < V + 2; v += NT) s_cstart[v] = 0;
    for (int i = tid; i < 2 * (R + 1); i += NT) s_bnd[i] = make_float2(kNeg, 0.f);
    if (tid == 0) {
        s_ll[0] = 0.f; s_ll[1] = 0.f; s_ll[2] = 0.f;
        for (int i = 0; i < NL + NS; ++i) mbar_init(bar_acts + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();
    if (want_grad)
        for (int i = tid; i < S; i += NT) atomicAdd(&s_cstart[s_lab[i] + 1], 1);
    __syncthreads();
    if (want_grad && w == 0) {  // exclusive scan of the class histogram
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

    using TrueT = std::integral_constant<bool, true>;
    using FalseT = std::integral_constant<bool, false>;

#ifdef CTC_B200_PROFILE
    // developer instrumentation: busy cycles of CTA 0 per role -> workspace header (u64 at +64):
    // [0] REC warp 0, [1] helper 0, [2] helper 1, [3] wall, [4] iterations, [5] T_b,
    // [6] helper 0 softmax part, [7] helper 0 gradient part
    unsigned long long* prof = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(p.status) + 64);
    const bool prof_on = blockIdx.x == 0 && lane == 0 && (w == 0 || w == R || w == R + 1);
    const int prof_slot = w == 0 ? 0 : (w == R ? 1 : 2);
    long long prof_t0 = clock64();
    const long long prof_start = prof_t0;
#define PROF_BEGIN() do { prof_t0 = clock64(); } while (0)
#define PROF_END() do { if (prof_on) atomicAdd(prof + prof_slot, (unsigned long long)(clock64() - prof_t0)); } while (0)
#define PROF_MARK(slot) do { if (prof_on && w == R) { atomicAdd(prof + (slot), (unsigned long long)(clock64() - prof_t1)); } prof_t1 = clock64(); } while (0)
    long long prof_t1 = prof_t0;
#else
#define PROF_BEGIN() do {} while (0)
#define PROF_END() do {} while (0)
#define PROF_MARK(slot) do {} while (0)
#endif

    if (is_rec) {
        // =============================================================================
        // REC
        // =============================================================================
        const int i0 = tid * P;
        int lab[P], pos[P];
        bool skip[P], selB[P], selY[P];
        float aB[P], aY[P];
        const int jB0 = S - i0;                            // partner pair index of my first blank
        const int pw_hi = max(jB0, 0) / (32 * P);          // at most two partner warps per thread
        const int pw_lo = max(pw_hi - 1, 0);
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const int i = i0 + k;
            lab[k] = s_lab[i];
            skip[k] = (i >= 1 && i < S && lab[k] != s_lab[i - 1]);
            pos[k] = i;                                    // padding pairs park their kNeg at slot i >= S
            if (want_grad && i < S) {                      // deterministic rank inside the class
                int r = 0;
                for (int j = 0; j < i; ++j) r += (s_lab[j] == lab[k]) ? 1 : 0;
                pos[k] = s_cstart[lab[k]] + r;
            }
            selB[k] = max(jB0 - k, 0) / (32 * P) == pw_hi;
            selY[k] = max(jB0 - 1 - k, 0) / (32 * P) == pw_hi;
            aB[k] = kNeg;
            aY[k] = kNeg;
        }
        if (tid == 0) aB[0] = 0.0f;  // virtual row "-1": log(1) in front of the first blank
        float off = 0.0f;            // this warp's exact integer offset (true = off + a)
        bool fresh = (w != 0);       // warp has not received any real value yet
        int par = 0;
        float ll_int = 0.f, ll_frac = 0.f;
        // thread / warp activity windows in sweep steps (widened by P pairs, see above)
        const unsigned win_store = (i0 - P <= S) ? (unsigned)(C + 3 * P) : 0u;   // (tt - i0 + P) < win
        const unsigned win_cons = (i0 <= S) ? (unsigned)(C + P) : 0u;            // (tt - i0) < win
        const int w_first = (32 * P * w - P <= S) ? 32 * P * w - P : 0x3fffffff;
        const int w_last = C + 32 * P * w + 32 * P - 1 + P;

        auto rec_step = [&](auto consume_tag, auto first_tag, int tt, const float* lp2,
                            const float* st, float* erow) {
            constexpr bool CONSUME = decltype(consume_tag)::value;
            constexpr bool FIRST = decltype(first_tag)::value;
            const bool wact = (tt >= w_first) && (tt <= w_last);
            float eBv[P], eYv[P];   // combined log2 occupancies (FIRST: fractional parts)
            float nBv[P], nYv[P];   // FIRST only: exact integer parts
            if (CONSUME) {
#pragma unroll
                for (int k = 0; k < P; ++k) { eBv[k] = kNeg; eYv[k] = kNeg; nBv[k] = 0.f; nYv[k] = 0.f; }
            }
            if (wact) {
                const float lpb = lp2[blank];
                float am1 = __shfl_up_sync(0xffffffffu, aY[P - 1], 1);
                float2 bq = make_float2(kNeg, 0.f);
                if (lane == 0) bq = s_bnd[par * (R + 1) + w];  // slot 0 is the constant (kNeg, 0)
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
                if (R > 1 && lane == 31) s_bnd[(par ^ 1) * (R + 1) + w + 1] = make_float2(aY[P - 1], off);

                if (!CONSUME) {
                    if (want_grad || tt == n_store - 1) {
                        float* row = lat_b + (size_t)(tbase + tsign * tt) * RS;
                        if ((unsigned)(tt - i0 + P) < win_store) {
                            store_vec<P>(row + i0, aB);
                            store_vec<P>(row + NP + i0, aY);
                        }
                        if (lane == 0) row[2 * NP + w] = off;
                    }
                } else if ((unsigned)(tt - i0) < win_cons) {
                    // a + partner - lp + (offsets - ll): the integer parts combine exactly
                    const float nhi = off + st[2 * NP + pw_hi], nlo = off + st[2 * NP + pw_lo];
                    if (FIRST) {
#pragma unroll
                        for (int k = 0; k < P; ++k) {
                            if (i0 + k <= S) { eBv[k] = (aB[k] + st[jB0 - k]) - lpb; nBv[k] = selB[k] ? nhi : nlo; }
                            if (i0 + k < S) { eYv[k] = (aY[k] + st[NP + jB0 - 1 - k]) - lpl[k]; nYv[k] = selY[k] ? nhi : nlo; }
                        }
                    } else {
                        const float chi = (nhi - ll_int) - ll_frac, clo = (nlo - ll_int) - ll_frac;
                        const float bhi = chi - lpb, blo = clo - lpb;
#pragma unroll
                        for (int k = 0; k < P; ++k) {
                            if (i0 + k <= S) eBv[k] = (aB[k] + st[jB0 - k]) + (selB[k] ? bhi : blo);
                            if (i0 + k < S) eYv[k] = (aY[k] + st[NP + jB0 - 1 - k]) + ((selY[k] ? chi : clo) - lpl[k]);
                        }
                    }
                }
            }
            if (CONSUME) {
                if (FIRST) {
                    // log-likelihood from the first combined row, ll = ll_int + ll_frac with an
                    // exact integer pivot: ll_int = max(n + rint(e)), ll_frac = log2 sum 2^(e + n - ll_int)
                    float pm = kNeg;
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        if (eBv[k] > kRealThresh) pm = fmaxf(pm, nBv[k] + rintf(eBv[k]));
                        if (eYv[k] > kRealThresh) pm = fmaxf(pm, nYv[k] + rintf(eYv[k]));
                    }
                    pm = warp_max(pm);
                    if (R > 1) {
                        if (lane == 0) s_red[w] = pm;
                        named_bar_sync(1, 32 * R);
                        for (int i = 0; i < R; ++i) pm = fmaxf(pm, s_red[i]);
                        named_bar_sync(1, 32 * R);
                    }
                    const bool infeasible = !(pm > kRealThresh);
                    ll_int = infeasible ? 0.f : pm;
                    float z = 0.f;
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        eBv[k] = (eBv[k] > kRealThresh) ? eBv[k] + (nBv[k] - ll_int) : kNeg;
                        eYv[k] = (eYv[k] > kRealThresh) ? eYv[k] + (nYv[k] - ll_int) : kNeg;
                        z += ex2f(eBv[k]) + ex2f(eYv[k]);
                    }
                    z = warp_sum(z);
                    if (R > 1) {
                        if (lane == 0) s_red[w] = z;
                        named_bar_sync(1, 32 * R);
                        z = 0.f;
                        for (int i = 0; i < R; ++i) z += s_red[i];
                        named_bar_sync(1, 32 * R);
                    }
                    ll_frac = infeasible ? 0.f : lg2f(z);
#pragma unroll
                    for (int k = 0; k < P; ++k) { eBv[k] -= ll_frac; eYv[k] -= ll_frac; }
                    if (tid == 0) {
                        s_ll[2] = infeasible ? 1.f : 0.f;
                        if (!rev) {
                            float out;
                            if (infeasible) out = p.zero_infinity ? 0.0f : CUDART_INF_F;
                            else out = (float)(-((double)ll_int + (double)ll_frac) * kLn2);
                            p.nll[b] = out;
                        }
                    }
                }
                if (want_grad) {
#pragma unroll
                    for (int k = 0; k < P; ++k) erow[NP + pos[k]] = eYv[k];
                    store_vec<P>(erow + i0, eBv);
                }
            }
            par ^= 1;
            if (R > 1) named_bar_sync(1, 32 * R);
        };

        Ring ring_lp2(NL), ring_part(NS);   // position of the chunk REC works on
        int e_buf = 0;
        for (int it = 0; it < nch + 2; ++it) {
            PROF_BEGIN();
            const int k = it - 1;
            if (k >= 0 && k < nch) {
                int tt0, rows;
                chunk_at(k, tt0, rows);
                const float* lp2c = s_lp2 + (size_t)ring_lp2.slot * TC * Vs;
                if (k < n1) {
                    for (int r = 0; r < rows; ++r)
                        rec_step(FalseT{}, FalseT{}, tt0 + r, lp2c + r * Vs, nullptr, nullptr);
                } else {
#ifdef CTC_B200_PROFILE
                    { long long tw = clock64();
#endif
                    mbar_wait(bar_part + ring_part.slot, ring_part.parity);   // TMA data landed
#ifdef CTC_B200_PROFILE
                      if (prof_on) atomicAdd(prof + 9, (unsigned long long)(clock64() - tw)); }
#endif
                    const float* stc = s_stage + (size_t)ring_part.slot * TC * RS;
                    float* ec = s_e + (size_t)e_buf * TC * ER;
                    int r = 0;
                    if (k == n1) { rec_step(TrueT{}, TrueT{}, tt0, lp2c, stc, ec); r = 1; }
                    if (want_grad)
                        for (; r < rows; ++r)
                            rec_step(TrueT{}, FalseT{}, tt0 + r, lp2c + r * Vs, stc + (size_t)r * RS,
                                     ec + (size_t)r * ER);
                    ring_part.advance();
                    e_buf ^= 1;
                }
                ring_lp2.advance();
            }
            PROF_END();
#ifdef CTC_B200_PROFILE
            if (prof_on) atomicAdd(prof + (it <= n1 ? 10 : 11), (unsigned long long)(clock64() - prof_t0));
#endif
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
        // ---- TMA issue: one lane per piece; piece = (row r, kind) ----------------------
        //   kind 0: logits row of chunk ka          -> lp2 ring slot
        //   kind 1/2/3: partner blank plane segment / label plane segment / warp offsets
        //               of chunk kp                  -> partner ring slot
        const int prow = lane >> 2, pkind = lane & 3;     // TC <= 8 rows x 4 kinds = 32 lanes
        auto issue_chunk = [&](int ka, int slot_a, int kp, int slot_p) {
            // ka / kp < 0: nothing of that kind to issue this time
            unsigned bytes = 0;
            const float* src = nullptr;
            float* dst = nullptr;
            if (pkind == 0) {
                if (ka >= 0) {
                    int tt0, rows;
                    chunk_at(ka, tt0, rows);
                    if (prow < rows) {
                        src = acts_b + (size_t)(tbase + tsign * (tt0 + prow)) * frame_stride;
                        dst = s_lp2 + ((size_t)slot_a * TC + prow) * Vs;
                        bytes = (unsigned)V * 4u;
                    }
                }
            } else if (kp >= 0) {
                int tt0, rows;
                chunk_at(kp, tt0, rows);
                if (prow < rows) {
                    const int tt = tt0 + prow;
                    // my in-band pairs (widened by P) <-> the partner pairs they map to
                    const int mlo = max(S - (Tb - tt) - P, 0), mhi = min(tt + P, S);
                    const int plo = max(S - 1 - mhi, 0) & ~3, phi = (S - mlo) | 3;
                    const float* srow = lat_b + (size_t)(tbase + tsign * tt) * RS;
                    float* drow = s_stage + ((size_t)slot_p * TC + prow) * RS;
                    if (pkind == 3) {
                        src = srow + 2 * NP; dst = drow + 2 * NP; bytes = (unsigned)(RS - 2 * NP) * 4u;
                    } else if (phi >= plo) {
                        const int o = (pkind == 2 ? NP : 0) + plo;
                        src = srow + o; dst = drow + o; bytes = (unsigned)(phi - plo + 1) * 4u;
                    }
                }
            }
            // expected transaction bytes per mbarrier, then the copies
            unsigned ba = pkind == 0 ? bytes : 0u, bp = pkind == 0 ? 0u : bytes;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ba += __shfl_xor_sync(0xffffffffu, ba, o);
                bp += __shfl_xor_sync(0xffffffffu, bp, o);
            }
            if (lane == 0) {
                if (ka >= 0) mbar_expect_tx(bar_acts + slot_a, ba);
                if (kp >= 0) mbar_expect_tx(bar_part + slot_p, bp);
            }
            __syncwarp();
            if (bytes) bulk_g2s(dst, src, bytes, pkind == 0 ? bar_acts + slot_a : bar_part + slot_p);
        };

        // ---- fused log_softmax of one staged row, in place (full warp) ----------------
        auto softmax_row = [&](float* row) {
            float4* row4 = reinterpret_cast<float4*>(row);
            if (V4 <= 32) {  // the whole row is one float4 per lane
                float4 q = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
                if (lane < V4) q = row4[lane];
                const float m = warp_max(fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
                q.x = (q.x - m) * kLog2e; q.y = (q.y - m) * kLog2e;
                q.z = (q.z - m) * kLog2e; q.w = (q.w - m) * kLog2e;
                const float lz = lg2f(warp_sum((ex2f(q.x) + ex2f(q.y)) + (ex2f(q.z) + ex2f(q.w))));
                if (lane < V4)
                    row4[lane] = make_float4(fmaxf(q.x - lz, kNeg), fmaxf(q.y - lz, kNeg),
                                             fmaxf(q.z - lz, kNeg), fmaxf(q.w - lz, kNeg));
            } else {
                float m = -CUDART_INF_F, z = 0.f;
                for (int c = lane; c < V4; c += 32) {
                    const float4 q = row4[c];
                    m = fmaxf(m, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
                }
                m = warp_max(m);
                for (int c = lane; c < V4; c += 32) {
                    float4 q = row4[c];
                    q.x = (q.x - m) * kLog2e; q.y = (q.y - m) * kLog2e;
                    q.z = (q.z - m) * kLog2e; q.w = (q.w - m) * kLog2e;
                    z += (ex2f(q.x) + ex2f(q.y)) + (ex2f(q.z) + ex2f(q.w));
                    row4[c] = q;
                }
                const float lz = lg2f(warp_sum(z));
                for (int c = lane; c < V4; c += 32) {
                    const float4 q = row4[c];
                    row4[c] = make_float4(fmaxf(q.x - lz, kNeg), fmaxf(q.y - lz, kNeg),
                                          fmaxf(q.z - lz, kNeg), fmaxf(q.w - lz, kNeg));
                }
            }
            if (lane == 0) row[V] = kNeg;  // what padding pairs gather
        };

        // ---- gradient row: occupancies -> class sums -> softmax - occupancy ----------
        //   blank cells: plain sum.  label cells are stored class-sorted, so the sum of
        //   class v is PS[cstart[v+1]] - PS[cstart[v]] of their exclusive prefix sums PS.
        auto grad_row = [&](float* erow, float* lp2row, float* g, bool infeasible) {
            if (infeasible) {
                const float fill = p.zero_infinity ? 0.0f : CUDART_NAN_F;
                for (int c = lane; c < V4; c += 32)
                    reinterpret_cast<float4*>(g)[c] = make_float4(fill, fill, fill, fill);
                return;
            }
            float4* eB4 = reinterpret_cast<float4*>(erow);
            float4* eY4 = reinterpret_cast<float4*>(erow + NP);
            float bs = 0.f, carry = 0.f;
            for (int c = lane; c * 4 <= S; c += 32) {           // blanks 0..S (padding holds kNeg)
                const float4 q = eB4[c];
                bs += (ex2f(q.x) + ex2f(q.y)) + (ex2f(q.z) + ex2f(q.w));
            }
            bs = warp_sum(bs);
            for (int base = 0; base < S; base += 128) {         // labels: 128 cells per round
                const int c = (base >> 2) + lane;
                float4 q = make_float4(kNeg, kNeg, kNeg, kNeg);
                if (c * 4 < NP) q = eY4[c];                     // slots in [S, NP) hold kNeg -> 0
                const float o0 = ex2f(q.x), o1 = o0 + ex2f(q.y), o2 = o1 + ex2f(q.z), o3 = o2 + ex2f(q.w);
                float inc = o3;                                 // inclusive scan of the lane totals
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += y;
                }
                const float ex = carry + (inc - o3);            // exclusive prefix of this lane
                if (c * 4 < NP) eY4[c] = make_float4(ex, ex + o0, ex + o1, ex + o2);   // PS[4c..4c+3]
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            __syncwarp();
            const float* PS = erow + NP;                        // PS[k] = sum of sorted cells < k
            for (int v = lane; v < V; v += 32) {
                const int k0 = s_cstart[v], k1 = s_cstart[v + 1];
                const float hi = (k1 < S) ? PS[k1] : carry;     // PS[S] = total
                const float lo = (k0 < S) ? PS[k0] : carry;
                const float occ = (hi - lo) + (v == blank ? bs : 0.f);
                lp2row[v] = gscale * (ex2f(lp2row[v]) - occ);
            }
            __syncwarp();
            for (int c = lane; c < V4; c += 32)                 // coalesced 128-bit row store
                reinterpret_cast<float4*>(g)[c] = reinterpret_cast<const float4*>(lp2row)[c];
        };

        // ---- the helper schedule ---------------------------------------------------------
        // Iteration `it`:  issue TMA { logits of chunk it+D+1, partner rows of chunk it+D };
        //                  softmax chunk it;  gradient rows of chunk it-2.
        // (REC runs chunk it-1.)  Partner rows of the first D+1 consume chunks cannot be
        // requested before the partner CTA wrote them: they are issued at the phase break.
        Ring iss_a(NL), iss_p(NS), sm_a(NL), gr_a(NL);
        int gr_e = 0;
        if (hw == 0) {
            for (int k = 0; k <= D; ++k) {            // prologue: logits of chunks 0..D
                if (k < nch) issue_chunk(k, iss_a.slot, -1, 0);
                iss_a.advance();
            }
        }
        for (int it = 0; it < nch + 2; ++it) {
            PROF_BEGIN();
            PROF_MARK(8);
            if (hw == 0) {
                const int ka = it + D + 1, kp = it + D;
                const bool do_a = ka < nch, do_p = want_grad && it >= n1 + 1 && kp < nch;
                if (do_a || do_p) issue_chunk(do_a ? ka : -1, iss_a.slot, do_p ? kp : -1, iss_p.slot);
                iss_a.advance();
                if (it >= n1 + 1) iss_p.advance();
            }
            if (it < nch) {                           // softmax of chunk `it`
                int tt0, rows;
                chunk_at(it, tt0, rows);
#ifdef CTC_B200_PROFILE
                { long long tw = clock64();
#endif
                mbar_wait(bar_acts + sm_a.slot, sm_a.parity);
#ifdef CTC_B200_PROFILE
                  if (prof_on && w == R) atomicAdd(prof + 12, (unsigned long long)(clock64() - tw)); }
#endif
#ifndef CTC_B200_NOHELP
                for (int r = hw; r < rows; r += H)
                    softmax_row(s_lp2 + ((size_t)sm_a.slot * TC + r) * Vs);
#endif
            }
            sm_a.advance();
            PROF_MARK(6);
            const int kg = it - 2;
            if (kg >= 0) {
                if (want_grad && kg >= n1 && kg < nch) {   // gradient rows of chunk it-2
                    int tt0, rows;
                    chunk_at(kg, tt0, rows);
                    const bool infeasible = s_ll[2] != 0.f;
#ifndef CTC_B200_NOHELP
                    for (int r = hw; r < rows; r += H)
#else
                    for (int r = hw; r < 0; r += H)
#endif
                        grad_row(s_e + ((size_t)gr_e * TC + r) * ER,
                                 s_lp2 + ((size_t)gr_a.slot * TC + r) * Vs,
                                 grad_b + (size_t)(tbase + tsign * (tt0 + r)) * frame_stride, infeasible);
                    gr_e ^= 1;
                    fence_proxy_async();   // my generic writes to the lp2 slot precede its next TMA fill
                }
                gr_a.advance();
            }
            PROF_MARK(7);
            PROF_END();
            __syncthreads();
            if (it == n1) {
                // Phase break: my REC warps have stored every row the partner will consume,
                // and (after the cluster barrier) vice versa.
                cluster_sync_all();
                if (hw == 0) {
                    fence_proxy_async();
                    for (int k = n1; k <= n1 + D; ++k) {
                        if (k < nch && (want_grad || k == n1)) issue_chunk(-1, 0, k, iss_p.slot);
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
