// ctc_abi.cu -- the extern "C" boundary declared in include/ctc_b200.h.
//
// Host-side launch logic for the kernels in ctc_kernels.cuh plus the
// host-buffer session.  No torch types here: this file builds into
// libctc_b200.so with nvcc alone and is what a non-Python caller links.
#include "ctc_kernels.cuh"
#include "ctc_pipe.cuh"
#include "ctc_lin.cuh"
#include "../../include/ctc_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

using namespace ctcb200;

namespace {

constexpr int kHeaderBytes = 256;          // workspace header: device status word
constexpr int kMaxSmemBytes = 200 * 1024;  // leave room under the 227 KB per-CTA limit
constexpr int kMinThreads = 32;
constexpr int kNumSmsHint = 148;           // B200; only shapes the shared-memory budget heuristic

thread_local cudaError_t g_last_cuda = cudaSuccess;

inline int cuda_fail(cudaError_t e) {
    g_last_cuda = e;
    return CTC_B200_CUDA_ERROR;
}
#define CTC_CUDA(call)                                  \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return cuda_fail(e__);  \
    } while (0)

// Programmatic dependent launch for the small kernels that follow the fused kernel (its fallback pass
// and the loss reduction): the launch is set up while the previous kernel drains; the kernels
// themselves wait for it with griddepcontrol.wait before their first global-memory read.
thread_local bool t_allow_pdl = true;   // the host-buffer session launches without it (measured slower there)

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                       Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

struct Geometry {
    int pipe;       // 2: linear-domain kernel (ctc_lin.cuh) with the log-domain pipe kernel as its
                    //    per-utterance fallback, 1: log-domain pipe kernel (ctc_pipe.cuh),
                    // 0: generic kernel (ctc_kernels.cuh)
    int P, NT, W, NP, chunk, RS, smem;
    int R, G, D;    // pipe only: recursion / gradient warps, fetch distance in chunks
    size_t lat_utt_stride;  // floats
    // pipe == 2: geometry of the linear kernel (the fields above describe the fallback)
    int lP, lNT, lNP, lchunk, lRS, lsmem, lR, lH, lD, lYS;
    size_t l_lat_utt_stride;
    size_t lattice_floats_per_utt() const { return pipe == 2 ? std::max(lat_utt_stride, l_lat_utt_stride) : lat_utt_stride; }
};

inline size_t flag_bytes(int n_utt) { return ((size_t)std::max(n_utt, 1) * 8 + 255) / 256 * 256; }
// softmax rows the linear kernel saves for the partner CTA's second half: [n_utt][T][V] fp32
inline size_t ysave_bytes(int T, int V, int n_utt) {
    return ((size_t)std::max(n_utt, 1) * (size_t)std::max(T, 1) * (size_t)V * 4 + 255) / 256 * 256;
}

int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

// Generic kernel: every warp does recursion + its share of softmax / gradient rows.
int pick_generic(int T, int V, int pairs, Geometry* g) {
    int P = pairs > 1024 ? (pairs > 2048 ? 4 : 2) : 1;
    const int q = env_int("CTC_B200_PAIRS", 0);  // tuning override (1, 2 or 4)
    if ((q == 1 || q == 2 || q == 4) && (pairs + q - 1) / q <= 1024) P = q;
    int NT = ((pairs + P - 1) / P + 31) / 32 * 32;
    NT = std::max(NT, P == 1 ? 128 : kMinThreads);
    g->pipe = 0;
    g->P = P;
    g->NT = NT;
    g->W = NT / 32;
    g->R = g->W;
    g->G = 0;
    g->D = 1;
    g->NP = NT * P;
    g->RS = 2 * g->NP + (g->W + 3) / 4 * 4;
    int chunk = kMaxChunk;
    for (;;) {
        SmemLayout lay(g->NP, g->W, V, chunk, g->RS);
        if (lay.total <= kMaxSmemBytes) {
            g->smem = lay.total;
            break;
        }
        if (chunk == 1) return CTC_B200_UNSUPPORTED;
        chunk >>= 1;
    }
    g->chunk = chunk;
    g->lat_utt_stride = (size_t)std::max(T, 1) * (size_t)g->RS;
    return CTC_B200_OK;
}

// Warp-specialised kernel: R recursion warps (P pairs per thread) + 1 load warp + G gradient warps.
bool pick_pipe(int T, int V, int pairs, int n_utt, Geometry* g) {
    if (V % 4) return false;   // TMA row copies need 16-byte aligned logit rows
    int P = pairs > 64 ? 4 : (pairs > 32 ? 2 : 1);
    const int q = env_int("CTC_B200_PAIRS", 0);
    if (q == 1 || q == 2 || q == 4) P = q;
    const int R = (pairs + 32 * P - 1) / (32 * P);
    int H = env_int("CTC_B200_HELPERS", 0);
    if (H < 1 || H > 8) H = pairs > 700 ? 4 : 2;
    const int NT = 32 * (R + H);
    if (NT > 1024) return false;
    // (chunk, fetch distance) candidates.  All CTAs should be co-resident (one wave: the
    // kernel is as long as its longest utterance), so the shared-memory budget per CTA is
    // an SM's 227 KB divided by the CTAs per SM the launch needs (at most 4).
    const int tc_env = env_int("CTC_B200_CHUNK", 0), d_env = env_int("CTC_B200_DIST", 0);
    const int cand[4][2] = {{4, 1}, {2, 2}, {2, 1}, {1, 1}};
    const int RS = 2 * 32 * P * R + (R + 3) / 4 * 4;
    for (int pass = 0; pass < 2; ++pass) {
        for (int ci = 0; ci < 4; ++ci) {
            int TC = cand[ci][0], D = cand[ci][1];
            if (tc_env == 1 || tc_env == 2 || tc_env == 4 || tc_env == 8) TC = tc_env;
            if (d_env >= 1 && d_env <= 4) D = d_env;
            PipeSmem lay(32 * P * R, R, V, TC, RS, D);
            const int need = std::max(1, std::min(4, (2 * std::max(n_utt, 1) + kNumSmsHint - 1) / kNumSmsHint));
            const int limit = pass == 0 ? std::min(kMaxSmemBytes, 227 * 1024 / need - 2048) : kMaxSmemBytes;
            if (lay.total > limit) continue;
            g->pipe = 1;
            g->P = P;
            g->NT = NT;
            g->W = NT / 32;
            g->R = R;
            g->G = H;
            g->D = D;
            g->NP = 32 * P * R;
            g->RS = RS;
            g->chunk = TC;
            g->smem = lay.total;
            g->lat_utt_stride = (size_t)std::max(T, 1) * (size_t)RS;
            return true;
        }
    }
    return false;
}

// Linear-domain kernel: R recursion warps with P pairs per thread in registers (P = 8 covers 256
// lattice slots per warp), R combine warps, H softmax / gradient warps.  Needs S_max + P <= 32 * P * R
// slots (alignment shift).
bool pick_lin(int T, int V, int S_max, int n_utt, Geometry* g) {
    if (V % 4) return false;
    // P = 8 always: one recursion warp covers 248 labels, and on B200 it is also the fastest choice for
    // short targets (P < 8 is kept for experiments: it trips the posterior-mass check more often)
    int P = 8;
    const int q = env_int("CTC_B200_LIN_PAIRS", 0);
    if (q == 1 || q == 2 || q == 4 || q == 8) P = q;
    int R = (S_max + P + 32 * P - 1) / (32 * P);
    if (R > 1 && P != 8) { P = 8; R = (S_max + P + 32 * P - 1) / (32 * P); }   // several warps: P = 8 only
    int H = env_int("CTC_B200_HELPERS", 0);
    // one helper warp (softmax, then gradient rows) when one recursion warp suffices: 4 warps per CTA
    // leave 128 registers per thread, which the two-rows-in-flight combine pass needs
    if (!(H == 1 || H == 2 || H == 4 || H == 8)) H = V > 256 ? 4 : (R == 1 ? 1 : 2);
    int NC = env_int("CTC_B200_COMB", 0);          // combine groups (warps per recursion warp)
    if (NC < 1 || NC > 4) NC = R <= 4 ? 2 : 1;   // two combine groups while the CTA stays within 512 threads
    int NT = 32 * ((1 + NC) * R + H);
    if (R == 1)   // the single-recursion-warp kernels are built for at most 256 threads
        while (NT > 256 && H > 1) { H >>= 1; NT = 32 * ((1 + NC) * R + H); }
    if (NT > 1024) return false;
    const int NP = 32 * P * R;
    const int RS = lin_row_stride(NP, P);
    const int YS = V <= 60 ? 80 : 0;   // fixed emission-ring row stride (an immediate in the kernel)
    const int tc_env = env_int("CTC_B200_CHUNK", 0);
    const int cand[3] = {4, 2, 1};
    for (int pass = 0; pass < 2; ++pass) {
        for (int ci = 0; ci < 3; ++ci) {
            int TC = cand[ci];
            if (tc_env >= 1 && tc_env <= 4) TC = tc_env;   // the kernel unrolls 4 rows
            const int YSc = TC == 4 ? YS : 0;   // the fixed-stride variants are built for chunks of 4 frames
            LinSmem lay(NP, R, V, TC, RS, YSc);
            const int need = std::max(1, std::min(4, (2 * std::max(n_utt, 1) + kNumSmsHint - 1) / kNumSmsHint));
            const int limit = pass == 0 ? std::min(kMaxSmemBytes, 227 * 1024 / need - 1024) : kMaxSmemBytes;
            if (lay.total > limit) continue;
            g->lP = P; g->lNT = NT; g->lNP = NP; g->lchunk = TC; g->lRS = RS; g->lsmem = lay.total;
            g->lR = R; g->lH = H; g->lD = NC; g->lYS = YSc;
            g->l_lat_utt_stride = (size_t)std::max(T, 1) * (size_t)RS;
            return true;
        }
    }
    return false;
}

int num_sms() {
    static int n = -1;
    if (n < 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            n = v;
        else
            n = 148;
    }
    return n;
}

int pick_geometry(int T, int V, int S_max, int n_utt, Geometry* g) {
    if (T < 0 || V < 1 || S_max < 0) return CTC_B200_INVALID_ARGUMENT;
    const int pairs = S_max + 1;
    if (pairs > 4096) return CTC_B200_UNSUPPORTED;  // targets longer than 4095 labels
    const char* force = std::getenv("CTC_B200_KERNEL");   // g: generic, p: log-domain pipe, default: linear
    const bool want_generic = force && force[0] == 'g';
    const bool want_pipe = force && force[0] == 'p';
    if (!want_generic && pick_pipe(T, V, pairs, n_utt, g)) {
        if (!want_pipe && pick_lin(T, V, S_max, n_utt, g)) g->pipe = 2;
        return CTC_B200_OK;
    }
    return pick_generic(T, V, pairs, g);
}

template <int P, int MAXT>
int launch_fused_pt(const FusedParams& prm, const Geometry& g, int n_utt, cudaStream_t st) {
    static int configured_smem = -1;  // per-process, per-instantiation high-water mark
    if (g.smem > configured_smem) {
        CTC_CUDA(cudaFuncSetAttribute(ctc_fused_kernel<P, MAXT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem));
        configured_smem = g.smem;
    }
    ctc_fused_kernel<P, MAXT><<<dim3(2 * n_utt), dim3(g.NT), g.smem, st>>>(prm);
    CTC_CUDA(cudaGetLastError());
    return CTC_B200_OK;
}

template <int P>
int launch_fused_p(const FusedParams& prm, const Geometry& g, int n_utt, cudaStream_t st) {
    // small CTAs get the full register file, large ones the 64-register cap
    return g.NT <= 256 ? launch_fused_pt<P, 256>(prm, g, n_utt, st)
                       : launch_fused_pt<P, 1024>(prm, g, n_utt, st);
}

template <int P, int MAXT, int MINB>
int launch_pipe_pt(const PipeParams& pp, const Geometry& g, int n_utt, cudaStream_t st) {
    static int configured_smem = -1;
    if (g.smem > configured_smem) {
        CTC_CUDA(cudaFuncSetAttribute(ctc_pipe_kernel<P, MAXT, MINB>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem));
        configured_smem = g.smem;
    }
    if (pp.redo != nullptr && t_allow_pdl && env_int("CTC_B200_PDL", 1))   // fallback pass right behind the linear kernel
        CTC_CUDA(launch_pdl(ctc_pipe_kernel<P, MAXT, MINB>, dim3(2 * n_utt), dim3(g.NT), (size_t)g.smem, st, pp));
    else
        ctc_pipe_kernel<P, MAXT, MINB><<<dim3(2 * n_utt), dim3(g.NT), g.smem, st>>>(pp);
    CTC_CUDA(cudaGetLastError());
    return CTC_B200_OK;
}

template <int P>
int launch_pipe_p(const PipeParams& pp, const Geometry& g, int n_utt, cudaStream_t st) {
    if (g.NT <= 128) return launch_pipe_pt<P, 128, 4>(pp, g, n_utt, st);
    if (g.NT <= 160) return launch_pipe_pt<P, 160, 4>(pp, g, n_utt, st);
    if (g.NT <= 256) return launch_pipe_pt<P, 256, 2>(pp, g, n_utt, st);
    if (g.NT <= 512) return launch_pipe_pt<P, 512, 1>(pp, g, n_utt, st);
    return launch_pipe_pt<P, 1024, 1>(pp, g, n_utt, st);
}

template <int P, int RC, int YS, int MAXT, int MINB, bool FIX = false>
int launch_lin_pt(const PipeParams& pp, int* flags, const Geometry& g, int n_utt, cudaStream_t st) {
    static int configured_smem = -1;
    if (g.lsmem > configured_smem) {
        CTC_CUDA(cudaFuncSetAttribute(ctc_lin_kernel<P, RC, YS, MAXT, MINB, FIX>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, g.lsmem));
        configured_smem = g.lsmem;
    }
    ctc_lin_kernel<P, RC, YS, MAXT, MINB, FIX><<<dim3(2 * n_utt), dim3(g.lNT), g.lsmem, st>>>(pp, flags);
    CTC_CUDA(cudaGetLastError());
    return CTC_B200_OK;
}

// One recursion warp (the common case): every stride of the kernel is a compile-time constant.
template <int P>
int launch_lin_r1(const PipeParams& pp, int* flags, const Geometry& g, int n_utt, cudaStream_t st) {
    if (g.lYS == 80) {
        if constexpr (P == 8) {   // the headline shape class: V = 48, REC + 2 x COMB + one helper warp
            if (g.lNT == 128 && g.lH == 1 && g.lD == 2 && pp.f.V == 48 && env_int("CTC_B200_NOFIX", 0) == 0)
                return launch_lin_pt<P, 1, 80, 128, 4, true>(pp, flags, g, n_utt, st);
        }
        if (g.lNT <= 128) return launch_lin_pt<P, 1, 80, 128, 4>(pp, flags, g, n_utt, st);
        if (g.lNT <= 160) return launch_lin_pt<P, 1, 80, 160, 4>(pp, flags, g, n_utt, st);
        return launch_lin_pt<P, 1, 80, 256, 2>(pp, flags, g, n_utt, st);
    }
    return g.lNT <= 128 ? launch_lin_pt<P, 1, 0, 128, 4>(pp, flags, g, n_utt, st)
                        : launch_lin_pt<P, 1, 0, 256, 2>(pp, flags, g, n_utt, st);
}

// Several recursion warps (targets longer than 248 labels): P = 8; compile-time strides for 2 and 4
// recursion warps (up to 1016 labels) with the V <= 60 emission ring, run-time strides otherwise.
int launch_lin_rn(const PipeParams& pp, int* flags, const Geometry& g, int n_utt, cudaStream_t st) {
    if (g.lYS == 80 && g.lNT <= 512) {
        if (g.lR == 2) return launch_lin_pt<8, 2, 80, 512, 1>(pp, flags, g, n_utt, st);
        if (g.lR == 4) return launch_lin_pt<8, 4, 80, 512, 1>(pp, flags, g, n_utt, st);
    }
    if (g.lNT <= 256) return launch_lin_pt<8, 0, 0, 256, 2>(pp, flags, g, n_utt, st);
    if (g.lNT <= 512) return launch_lin_pt<8, 0, 0, 512, 1>(pp, flags, g, n_utt, st);
    return launch_lin_pt<8, 0, 0, 1024, 1>(pp, flags, g, n_utt, st);
}

int launch_fused(const float* acts, const int32_t* targets, const int32_t* tgt_offsets,
                 const int32_t* in_lens, const int32_t* tgt_lens, int T, int N, int V, int S_max,
                 int blank, int zero_infinity, int utt_begin, int utt_count, float* nll,
                 float* grad, const float* grad_scale, float* lattice, size_t lattice_bytes,
                 int* status_word, int* flags, cudaStream_t st) {
    if (!acts || !targets || !tgt_offsets || !in_lens || !tgt_lens || !nll || !status_word)
        return CTC_B200_INVALID_ARGUMENT;
    if (T < 0 || N < 0 || V < 1 || blank < 0 || blank >= V || utt_begin < 0 || utt_count < 0 ||
        utt_begin + utt_count > N)
        return CTC_B200_INVALID_ARGUMENT;
    if (utt_count == 0) return CTC_B200_OK;
    Geometry g;
    int rc = pick_geometry(T, V, S_max, N, &g);
    if (rc != CTC_B200_OK) return rc;
    const size_t lat_need = g.lattice_floats_per_utt() * sizeof(float) * (size_t)utt_count;
    const size_t ys_need = 0;   // (saved softmax rows: tried, no faster; the kernel keeps the hook)
    if (!lattice || lattice_bytes < lat_need + ys_need) return CTC_B200_WORKSPACE_TOO_SMALL;
    if (g.pipe == 2 && !flags) return CTC_B200_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(lattice) & 15) || (reinterpret_cast<uintptr_t>(acts) & 15) ||
        (grad && (reinterpret_cast<uintptr_t>(grad) & 15)))
        return CTC_B200_INVALID_ARGUMENT;

    FusedParams prm;
    prm.acts = acts;
    prm.targets = targets;
    prm.tgt_off = tgt_offsets;
    prm.in_lens = in_lens;
    prm.tgt_lens = tgt_lens;
    prm.grad_scale = grad_scale;
    prm.nll = nll;
    prm.grad = grad;
    prm.lattice = lattice;
    prm.ysave = ys_need ? reinterpret_cast<float*>(reinterpret_cast<char*>(lattice) + (lat_need + 255) / 256 * 256) : nullptr;
    prm.status = status_word;
    prm.lat_utt_stride = (long long)g.lat_utt_stride;
    prm.T = T;
    prm.N = N;
    prm.V = V;
    prm.blank = blank;
    prm.zero_infinity = zero_infinity;
    prm.utt_begin = utt_begin;
    prm.row_stride = g.RS;
    prm.chunk = g.chunk;
    if (g.pipe == 2) {
        // linear-domain kernel; `flags` ([2 * utt_count], indexed from utt_begin) marks the utterances
        // whose posterior-mass check failed, which the log-domain kernel below then recomputes
        PipeParams lp;
        lp.f = prm;
        lp.f.lat_utt_stride = (long long)g.l_lat_utt_stride;
        lp.f.row_stride = g.lRS;
        lp.f.chunk = g.lchunk;
        lp.R = g.lR;
        lp.H = g.lH;
        lp.NP = g.lNP;
        lp.D = g.lD;
        // co-resident CTAs rotate their warp roles only when the CTA has a multiple of 4 warps (else
        // consecutive CTAs already start on different SM sub-partitions)
        lp.rotate = (env_int("CTC_B200_ROTATE", 1) && (g.lNT / 32) % 4 == 0) ? num_sms() : 0;
        lp.redo = nullptr;
        // CTA placement for a launch that fits one wave with an incomplete last "layer" (C2: 256 clusters on
        // 74 SM pairs = 3 full layers + 34 clusters): the batch is sorted by length (dataloader.py:53), so with
        // the identity mapping the LONGEST utterances share their SMs with a fourth CTA.  Rotating the
        // utterances by whole layers moves them to the SM pairs that host only three (measured on B200, C2:
        // 0.263 ms -> 0.248 ms with a rotation of two layers; rotations that are not whole layers: 0.27 ms).
        {
            const int pairs = std::max(1, num_sms() / 2), layers = utt_count / pairs;
            int rot = (utt_count > pairs && utt_count <= 4 * pairs && utt_count % pairs != 0)
                          ? pairs * std::max(1, layers - 1) : 0;
            const int e = env_int("CTC_B200_UTT_ROT", -1);
            if (e >= 0) rot = e;
            lp.utt_rot = std::max(0, std::min(rot, utt_count - 1));
        }
        int* fl = flags - 2 * (ptrdiff_t)utt_begin;   // kernels index flags by absolute utterance
        if (g.lR == 1) {
            switch (g.lP) {
                case 1: rc = launch_lin_r1<1>(lp, fl, g, utt_count, st); break;
                case 2: rc = launch_lin_r1<2>(lp, fl, g, utt_count, st); break;
                case 4: rc = launch_lin_r1<4>(lp, fl, g, utt_count, st); break;
                case 8: rc = launch_lin_r1<8>(lp, fl, g, utt_count, st); break;
                default: rc = CTC_B200_UNSUPPORTED;
            }
        } else {
            rc = g.lP == 8 ? launch_lin_rn(lp, fl, g, utt_count, st) : CTC_B200_UNSUPPORTED;
        }
        if (rc != CTC_B200_OK) return rc;
#ifdef CTC_B200_DEV_KNOBS   // developer builds only: look at the linear kernel's own output for flagged utterances
        if (env_int("CTC_B200_NOFALLBACK", 0)) return CTC_B200_OK;
#endif
    }
    if (g.pipe) {
        PipeParams pp;
        pp.f = prm;
        pp.redo = g.pipe == 2 ? flags - 2 * (ptrdiff_t)utt_begin : nullptr;
        pp.utt_rot = 0;
        pp.R = g.R;
        pp.H = g.G;
        pp.NP = g.NP;
        pp.D = g.D;
        pp.rotate = env_int("CTC_B200_ROTATE", 1) ? num_sms() : 0;
        switch (g.P) {
            case 1: return launch_pipe_p<1>(pp, g, utt_count, st);
            case 2: return launch_pipe_p<2>(pp, g, utt_count, st);
            case 4: return launch_pipe_p<4>(pp, g, utt_count, st);
        }
        return CTC_B200_UNSUPPORTED;
    }
    switch (g.P) {
        case 1: return launch_fused_p<1>(prm, g, utt_count, st);
        case 2: return launch_fused_p<2>(prm, g, utt_count, st);
        case 4: return launch_fused_p<4>(prm, g, utt_count, st);
    }
    return CTC_B200_UNSUPPORTED;
}

int status_from_bits(int bits) {
    if (bits & kStatusBadLabel) return CTC_B200_BAD_LABEL;
    if (bits & kStatusBadLength) return CTC_B200_BAD_LENGTH;
    if (bits & kStatusPeerTimeout) return CTC_B200_PEER_TIMEOUT;
    return CTC_B200_OK;
}

}  // namespace

extern "C" {

int ctc_b200_version(void) { return 1000; }

const char* ctc_b200_status_string(int status) {
    switch (status) {
        case CTC_B200_OK: return "ok";
        case CTC_B200_INVALID_ARGUMENT: return "invalid argument";
        case CTC_B200_WORKSPACE_TOO_SMALL: return "workspace too small";
        case CTC_B200_UNSUPPORTED: return "unsupported problem size";
        case CTC_B200_CUDA_ERROR: return "CUDA error";
        case CTC_B200_BAD_LABEL: return "target label outside [0, V)";
        case CTC_B200_BAD_LENGTH: return "input length > T or target length > S_max";
        case CTC_B200_PEER_TIMEOUT: return "fused loss all-reduce: a peer did not arrive";
    }
    return "unknown status";
}

const char* ctc_b200_last_cuda_error(void) { return cudaGetErrorString(g_last_cuda); }

int ctc_b200_get_geometry(int T, int n_utt, int V, int S_max, ctc_b200_geometry* out) {
    if (!out || n_utt < 0) return CTC_B200_INVALID_ARGUMENT;
    Geometry g;
    int rc = pick_geometry(T, V, S_max, n_utt, &g);
    if (rc != CTC_B200_OK) return rc;
    const bool lin = g.pipe == 2;
    out->kernel = g.pipe;
    out->rec_warps = lin ? g.lR : g.R;
    out->grad_warps = lin ? g.lH : g.G;
    out->pairs_per_thread = lin ? g.lP : g.P;
    out->threads = lin ? g.lNT : g.NT;
    out->chunk = lin ? g.lchunk : g.chunk;
    out->row_stride = lin ? g.lRS : g.RS;
    out->smem_bytes = lin ? g.lsmem : g.smem;
    // [256 B header: status word][redo flags: 2 ints per utterance][lattice]
    out->workspace_bytes = kHeaderBytes + flag_bytes(n_utt) +
                           (g.lattice_floats_per_utt() * sizeof(float) * (size_t)n_utt + 255) / 256 * 256 +
                           0;
    return CTC_B200_OK;
}

int ctc_b200_workspace_bytes(int T, int N, int V, int S_max, size_t* bytes) {
    if (!bytes) return CTC_B200_INVALID_ARGUMENT;
    ctc_b200_geometry g;
    int rc = ctc_b200_get_geometry(T, N, V, S_max, &g);
    if (rc == CTC_B200_OK) *bytes = g.workspace_bytes;
    return rc;
}

int ctc_b200_fwd_bwd_range_f32(const float* acts, const int32_t* targets,
                               const int32_t* tgt_offsets, const int32_t* in_lens,
                               const int32_t* tgt_lens, int T, int N, int V, int S_max,
                               int blank, int zero_infinity, int utt_begin, int utt_count,
                               float* nll, float* grad, const float* grad_scale,
                               void* workspace, size_t workspace_bytes, void* stream) {
    const size_t head = (size_t)kHeaderBytes + flag_bytes(utt_count);
    if (!workspace || workspace_bytes < head) return CTC_B200_WORKSPACE_TOO_SMALL;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return CTC_B200_INVALID_ARGUMENT;
    char* ws = static_cast<char*>(workspace);
    return launch_fused(acts, targets, tgt_offsets, in_lens, tgt_lens, T, N, V, S_max, blank,
                        zero_infinity, utt_begin, utt_count, nll, grad, grad_scale,
                        reinterpret_cast<float*>(ws + head), workspace_bytes - head,
                        reinterpret_cast<int*>(ws), reinterpret_cast<int*>(ws + kHeaderBytes),
                        static_cast<cudaStream_t>(stream));
}

int ctc_b200_fwd_bwd_f32(const float* acts, const int32_t* targets, const int32_t* tgt_offsets,
                         const int32_t* in_lens, const int32_t* tgt_lens, int T, int N, int V,
                         int S_max, int blank, int zero_infinity, float* nll, float* grad,
                         const float* grad_scale, void* workspace, size_t workspace_bytes,
                         void* stream) {
    return ctc_b200_fwd_bwd_range_f32(acts, targets, tgt_offsets, in_lens, tgt_lens, T, N, V,
                                      S_max, blank, zero_infinity, 0, N, nll, grad, grad_scale,
                                      workspace, workspace_bytes, stream);
}

int ctc_b200_scale_grad_f32(float* grad, const float* scale, int per_utt, int T, int N, int V,
                            void* stream) {
    if (!grad || !scale || T < 0 || N < 0 || V < 1) return CTC_B200_INVALID_ARGUMENT;
    if (T == 0 || N == 0) return CTC_B200_OK;
    const int threads = 256;
    const size_t per_utt_elems = (size_t)T * V;
    int gy = (int)std::min<size_t>((per_utt_elems + threads * 8 - 1) / (threads * 8), 64);
    gy = std::max(gy, 1);
    ctc_scale_grad_kernel<<<dim3(N, gy), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        grad, scale, per_utt, T, N, V);
    CTC_CUDA(cudaGetLastError());
    return CTC_B200_OK;
}

int ctc_b200_reduce_loss_f32(const float* nll, const int32_t* tgt_lens, int N, int reduction,
                             float* out2, float* loss, void* stream) {
    if (!nll || !tgt_lens || !out2 || N < 0) return CTC_B200_INVALID_ARGUMENT;
    if (t_allow_pdl && env_int("CTC_B200_PDL", 1))
        CTC_CUDA(launch_pdl(ctc_reduce_loss_kernel, dim3(1), dim3(256), (size_t)0, static_cast<cudaStream_t>(stream),
                            nll, tgt_lens, N, reduction == CTC_B200_REDUCE_MEAN ? 1 : 2, out2, loss));
    else
        ctc_reduce_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            nll, tgt_lens, N, reduction == CTC_B200_REDUCE_MEAN ? 1 : 2, out2, loss);
    CTC_CUDA(cudaGetLastError());
    return CTC_B200_OK;
}

static int launch_loss_allreduce(const float* nll, const int32_t* tgt_lens, int N, int reduction,
                                 const float* pair_in, void* const* peer_bufs, int rank, int world_size,
                                 unsigned seq, float* out2, float* loss, void* workspace, void* stream) {
    if ((!pair_in && (!nll || !tgt_lens)) || !out2 || !peer_bufs || !workspace || N < 0 || world_size < 1 ||
        world_size > CTC_B200_MAX_PEERS || rank < 0 || rank >= world_size || seq == 0)
        return CTC_B200_INVALID_ARGUMENT;
    PeerBufs pb;
    for (int r = 0; r < CTC_B200_MAX_PEERS; ++r) {
        void* q = r < world_size ? peer_bufs[r] : nullptr;
        if (r < world_size && (!q || (reinterpret_cast<uintptr_t>(q) & 15))) return CTC_B200_INVALID_ARGUMENT;
        pb.p[r] = static_cast<float4*>(q);
    }
    const int mode = reduction == CTC_B200_REDUCE_MEAN ? 1 : 2;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (t_allow_pdl && env_int("CTC_B200_PDL", 1))
        CTC_CUDA(launch_pdl(ctc_reduce_loss_allreduce_kernel, dim3(1), dim3(256), (size_t)0, st, nll, tgt_lens,
                            N, mode, pair_in, pb, rank, world_size, seq, out2, loss, static_cast<int*>(workspace)));
    else
        ctc_reduce_loss_allreduce_kernel<<<1, 256, 0, st>>>(nll, tgt_lens, N, mode, pair_in, pb, rank, world_size,
                                                            seq, out2, loss, static_cast<int*>(workspace));
    CTC_CUDA(cudaGetLastError());
    return CTC_B200_OK;
}

int ctc_b200_reduce_loss_allreduce_f32(const float* nll, const int32_t* tgt_lens, int N, int reduction,
                                       void* const* peer_bufs, int rank, int world_size, unsigned seq,
                                       float* out2, float* loss, void* workspace, void* stream) {
    return launch_loss_allreduce(nll, tgt_lens, N, reduction, nullptr, peer_bufs, rank, world_size, seq, out2,
                                 loss, workspace, stream);
}

int ctc_b200_allreduce_pair_f32(float* out2, int reduction, void* const* peer_bufs, int rank, int world_size,
                                unsigned seq, float* loss, void* status_word, void* stream) {
    return launch_loss_allreduce(nullptr, nullptr, 0, reduction, out2, peer_bufs, rank, world_size, seq, out2,
                                 loss, status_word, stream);
}

int ctc_b200_check_status(const void* workspace, void* stream) {
    if (!workspace) return CTC_B200_INVALID_ARGUMENT;
    int bits = 0;
    CTC_CUDA(cudaMemcpyAsync(&bits, workspace, sizeof(int), cudaMemcpyDeviceToHost,
                             static_cast<cudaStream_t>(stream)));
    CTC_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return status_from_bits(bits);
}

int ctc_b200_clear_status(void* workspace, void* stream) {
    if (!workspace) return CTC_B200_INVALID_ARGUMENT;
    CTC_CUDA(cudaMemsetAsync(workspace, 0, kHeaderBytes, static_cast<cudaStream_t>(stream)));
    return CTC_B200_OK;
}

// ---------------------------------------------------------------------------
// Host-buffer session
// ---------------------------------------------------------------------------
struct ctc_b200_session {
    int T, N, V, S_max, max_targets, n_slices;
    Geometry geo;
    float* d_acts = nullptr;
    float* d_grad = nullptr;
    float* d_lattice = nullptr;
    size_t lattice_bytes = 0;
    int* d_flags = nullptr;    // [2 * N] redo flags of the linear kernel
    char* d_small = nullptr;   // [targets | tgt_off | in_lens | tgt_lens | scale]
    char* h_small = nullptr;   // pinned mirror
    size_t small_bytes = 0;
    char* d_res = nullptr;     // [out2 (2 f32) | status (i32) | pad | nll (N f32)]
    char* h_res = nullptr;     // pinned mirror
    size_t res_bytes = 0;
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    std::vector<cudaEvent_t> ev;
    // one compute stream per slice: a slice's kernels start when ITS copy has landed, whether or not the
    // previous slice's (latency-bound) kernel has finished; s_comp joins them for the loss reduction
    std::vector<cudaStream_t> s_slice;
    std::vector<cudaEvent_t> ev_done;
    int last_launches = 0;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int ctc_b200_session_create(int T, int N, int V, int S_max, int max_targets, int n_slices,
                            ctc_b200_session** out) {
    if (!out || T < 1 || N < 1 || V < 1 || S_max < 0 || max_targets < 0)
        return CTC_B200_INVALID_ARGUMENT;
    ctc_b200_session* s = new (std::nothrow) ctc_b200_session();
    if (!s) return CTC_B200_INVALID_ARGUMENT;
    s->T = T; s->N = N; s->V = V; s->S_max = S_max; s->max_targets = max_targets;
    s->n_slices = std::max(1, std::min(n_slices, N));
    int rc = pick_geometry(T, V, S_max, N, &s->geo);
    if (rc != CTC_B200_OK) { delete s; return rc; }
    const size_t nact = (size_t)T * N * V * sizeof(float);
    // (slices run one after the other, so the saved-softmax region of a slice may overlap the lattice of
    // later slices; the allocation only has to leave room behind the last slice)
    s->lattice_bytes = s->geo.lattice_floats_per_utt() * sizeof(float) * (size_t)N +
                       0;
    s->small_bytes = align_up((size_t)std::max(max_targets, 1) * 4, 16) + 4 * align_up((size_t)N * 4, 16);
    s->res_bytes = 16 + (size_t)N * 4;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    ok(cudaMalloc(&s->d_acts, nact));
    ok(cudaMalloc(&s->d_grad, nact));
    ok(cudaMalloc(&s->d_lattice, s->lattice_bytes));
    ok(cudaMalloc(&s->d_flags, flag_bytes(N)));
    ok(cudaMalloc(&s->d_small, s->small_bytes));
    ok(cudaMalloc(&s->d_res, s->res_bytes));
    ok(cudaMallocHost(&s->h_small, s->small_bytes));
    ok(cudaMallocHost(&s->h_res, s->res_bytes));
    ok(cudaStreamCreateWithFlags(&s->s_copy, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&s->s_comp, cudaStreamNonBlocking));
    s->ev.resize(s->n_slices + 1, nullptr);
    for (auto& ev : s->ev) ok(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    s->s_slice.resize(s->n_slices, nullptr);
    s->ev_done.resize(s->n_slices, nullptr);
    for (auto& st : s->s_slice) ok(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto& ev : s->ev_done) ok(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (e != cudaSuccess) {
        ctc_b200_session_destroy(s);
        return cuda_fail(e);
    }
    *out = s;
    return CTC_B200_OK;
}

int ctc_b200_session_destroy(ctc_b200_session* s) {
    if (!s) return CTC_B200_OK;
    for (auto ev : s->ev) if (ev) cudaEventDestroy(ev);
    for (auto ev : s->ev_done) if (ev) cudaEventDestroy(ev);
    for (auto st : s->s_slice) if (st) cudaStreamDestroy(st);
    if (s->s_copy) cudaStreamDestroy(s->s_copy);
    if (s->s_comp) cudaStreamDestroy(s->s_comp);
    cudaFree(s->d_acts); cudaFree(s->d_grad); cudaFree(s->d_lattice); cudaFree(s->d_flags);
    cudaFree(s->d_small); cudaFree(s->d_res);
    if (s->h_small) cudaFreeHost(s->h_small);
    if (s->h_res) cudaFreeHost(s->h_res);
    delete s;
    return CTC_B200_OK;
}

float* ctc_b200_session_grad_device(ctc_b200_session* s) { return s ? s->d_grad : nullptr; }
int ctc_b200_session_last_launches(const ctc_b200_session* s) { return s ? s->last_launches : 0; }

int ctc_b200_session_run_host_f32(ctc_b200_session* s, const float* acts_host,
                                  const int32_t* targets_host, int n_targets,
                                  const int32_t* in_lens_host, const int32_t* tgt_lens_host,
                                  int blank, int reduction, int zero_infinity, int want_grad,
                                  float* loss_host, float* nll_host, float* grad_host) {
    if (!s || !acts_host || (!targets_host && n_targets > 0) || !in_lens_host || !tgt_lens_host ||
        !loss_host || n_targets < 0 || n_targets > s->max_targets)
        return CTC_B200_INVALID_ARGUMENT;
    const int T = s->T, N = s->N, V = s->V;
    // ---- host prep into pinned staging: offsets, lengths, per-utterance scale ---
    const size_t o_tg = 0;
    const size_t o_off = align_up((size_t)std::max(s->max_targets, 1) * 4, 16);
    const size_t o_il = o_off + align_up((size_t)N * 4, 16);
    const size_t o_tl = o_il + align_up((size_t)N * 4, 16);
    const size_t o_sc = o_tl + align_up((size_t)N * 4, 16);
    int32_t* h_tg = reinterpret_cast<int32_t*>(s->h_small + o_tg);
    int32_t* h_off = reinterpret_cast<int32_t*>(s->h_small + o_off);
    int32_t* h_il = reinterpret_cast<int32_t*>(s->h_small + o_il);
    int32_t* h_tl = reinterpret_cast<int32_t*>(s->h_small + o_tl);
    float* h_sc = reinterpret_cast<float*>(s->h_small + o_sc);
    long long acc = 0;
    for (int b = 0; b < N; ++b) {
        const int S = tgt_lens_host[b], Tb = in_lens_host[b];
        if (S < 0 || S > s->S_max || Tb < 0 || Tb > T) return CTC_B200_BAD_LENGTH;
        h_off[b] = (int32_t)acc;
        acc += S;
        h_il[b] = Tb;
        h_tl[b] = S;
        h_sc[b] = (reduction == CTC_B200_REDUCE_MEAN) ? 1.0f / ((float)N * (float)std::max(S, 1)) : 1.0f;
    }
    if (acc != n_targets) return CTC_B200_INVALID_ARGUMENT;
    if (n_targets) std::memcpy(h_tg, targets_host, (size_t)n_targets * 4);
    CTC_CUDA(cudaMemcpyAsync(s->d_small, s->h_small, s->small_bytes, cudaMemcpyHostToDevice, s->s_copy));
    CTC_CUDA(cudaMemsetAsync(s->d_res, 0, 16, s->s_copy));
    CTC_CUDA(cudaEventRecord(s->ev[s->n_slices], s->s_copy));
    CTC_CUDA(cudaStreamWaitEvent(s->s_comp, s->ev[s->n_slices], 0));

    const int32_t* d_tg = reinterpret_cast<const int32_t*>(s->d_small + o_tg);
    const int32_t* d_off = reinterpret_cast<const int32_t*>(s->d_small + o_off);
    const int32_t* d_il = reinterpret_cast<const int32_t*>(s->d_small + o_il);
    const int32_t* d_tl = reinterpret_cast<const int32_t*>(s->d_small + o_tl);
    const float* d_sc = reinterpret_cast<const float*>(s->d_small + o_sc);
    float* d_out2 = reinterpret_cast<float*>(s->d_res);
    int* d_status = reinterpret_cast<int*>(s->d_res + 8);
    float* d_nll = reinterpret_cast<float*>(s->d_res + 16);

    // ---- slice pipeline: H2D of slice k+1 overlaps the kernel of slice k -------
    int launches = 0;
    const size_t pitch = (size_t)N * V * sizeof(float);
    for (int k = 0; k < s->n_slices; ++k) {
        const int b0 = (int)((long long)N * k / s->n_slices);
        const int b1 = (int)((long long)N * (k + 1) / s->n_slices);
        if (b1 == b0) continue;
        CTC_CUDA(cudaMemcpy2DAsync(s->d_acts + (size_t)b0 * V, pitch, acts_host + (size_t)b0 * V,
                                   pitch, (size_t)(b1 - b0) * V * sizeof(float), (size_t)T,
                                   cudaMemcpyHostToDevice, s->s_copy));
        CTC_CUDA(cudaEventRecord(s->ev[k], s->s_copy));
        cudaStream_t sk = env_int("CTC_B200_SLICE_STREAMS", 1) ? s->s_slice[k] : s->s_comp;
        CTC_CUDA(cudaStreamWaitEvent(sk, s->ev[s->n_slices], 0));   // targets / lengths / cleared status
        CTC_CUDA(cudaStreamWaitEvent(sk, s->ev[k], 0));              // this slice's logits
        t_allow_pdl = false;
        int rc = launch_fused(s->d_acts, d_tg, d_off, d_il, d_tl, T, N, V, s->S_max, blank,
                              zero_infinity, b0, b1 - b0, d_nll, want_grad ? s->d_grad : nullptr,
                              d_sc, s->d_lattice + (size_t)b0 * s->geo.lattice_floats_per_utt(),
                              s->lattice_bytes - (size_t)b0 * s->geo.lattice_floats_per_utt() * sizeof(float),
                              d_status, s->d_flags + 2 * (size_t)b0, sk);
        t_allow_pdl = true;
        if (rc != CTC_B200_OK) return rc;
        CTC_CUDA(cudaEventRecord(s->ev_done[k], sk));
        CTC_CUDA(cudaStreamWaitEvent(s->s_comp, s->ev_done[k], 0));
        launches += s->geo.pipe == 2 ? 2 : 1;
    }
    ctc_reduce_loss_kernel<<<1, 256, 0, s->s_comp>>>(
        d_nll, d_tl, N, reduction == CTC_B200_REDUCE_MEAN ? 1 : 2, d_out2, nullptr);
    CTC_CUDA(cudaGetLastError());
    ++launches;
    CTC_CUDA(cudaMemcpyAsync(s->h_res, s->d_res, nll_host ? s->res_bytes : 16,
                             cudaMemcpyDeviceToHost, s->s_comp));
    if (grad_host && want_grad)
        CTC_CUDA(cudaMemcpyAsync(grad_host, s->d_grad, (size_t)T * N * V * sizeof(float),
                                 cudaMemcpyDeviceToHost, s->s_comp));
    CTC_CUDA(cudaStreamSynchronize(s->s_comp));
    s->last_launches = launches;

    const float* h_out2 = reinterpret_cast<const float*>(s->h_res);
    const int bits = *reinterpret_cast<const int*>(s->h_res + 8);
    *loss_host = (reduction == CTC_B200_REDUCE_MEAN) ? h_out2[0] / (float)N : h_out2[0];
    if (nll_host) std::memcpy(nll_host, s->h_res + 16, (size_t)N * 4);
    return status_from_bits(bits);
}

}  // extern "C"
