// ctc_abi.cu -- the extern "C" boundary declared in include/ctc_b200.h.
//
// Host-side geometry choice and launch logic, the small kernels (ctc_small.cuh) and the
// host-buffer session.  No torch types here: libctc_b200.so is built with nvcc alone
// (ctc_abi.cu + ctc_launch_lin.cu + ctc_launch_log.cu + ctc_decode.cu) and is what a non-Python caller links.
#include "ctc_launch.h"
#include "ctc_small.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

using namespace ctcb200;

namespace ctcb200 {

thread_local cudaError_t g_last_cuda = cudaSuccess;

cudaError_t last_cuda_error_set(cudaError_t e) {
    if (e != cudaSuccess) g_last_cuda = e;
    return e;
}

const Env& env() {
    static const Env e;   // thread-safe one-time initialisation
    return e;
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return dev;
}

int num_sms() {
    static std::atomic<int> n[kMaxDevices];
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices) return 148;
    int v = n[dev].load(std::memory_order_relaxed);
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
            cudaGetLastError();
            return 148;   // B200 (no device to ask: the geometry query still answers)
        }
        n[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

cudaError_t ensure_smem(const void* func, SmemMark& mark, int bytes) {
    const int dev = current_device();
    if (dev >= 0 && dev < kMaxDevices && mark.v[dev].load(std::memory_order_acquire) >= bytes) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return last_cuda_error_set(e);
    if (dev >= 0 && dev < kMaxDevices) {
        int old = mark.v[dev].load(std::memory_order_relaxed);
        while (old < bytes && !mark.v[dev].compare_exchange_weak(old, bytes, std::memory_order_release)) {}
    }
    return cudaSuccess;
}

}  // namespace ctcb200

namespace {

constexpr int kHeaderBytes = 256;          // workspace header: [0] status word, [16] utterance queue (2 ints)
constexpr int kQueueOffset = 16;
constexpr int kMaxSmemBytes = 200 * 1024;  // leave room under the 227 KB per-CTA limit
constexpr int kMinThreads = 32;
constexpr int kNumSmsHint = 148;           // B200; only shapes the shared-memory budget heuristic

inline int cuda_fail(cudaError_t e) {
    last_cuda_error_set(e);
    return CTC_B200_CUDA_ERROR;
}
#define CTC_CUDA(call)                                  \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return cuda_fail(e__);  \
    } while (0)

thread_local bool t_allow_pdl = true;   // the host-buffer session launches without it (measured slower there)

inline size_t flag_bytes(int n_utt) { return ((size_t)std::max(n_utt, 1) * 8 + 255) / 256 * 256; }

// Generic kernel: every warp does recursion + its share of softmax / gradient rows.
int pick_generic(int T, int V, int pairs, Geometry* g) {
    const int P = pairs > 1024 ? (pairs > 2048 ? 4 : 2) : 1;
    int NT = ((pairs + P - 1) / P + 31) / 32 * 32;
    NT = std::max(NT, P == 1 ? 128 : kMinThreads);
    g->base = 0;
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

// Warp-specialised log-domain kernel: R recursion warps (P pairs per thread) + H helper warps.
bool pick_pipe(int T, int V, int pairs, int n_utt, Geometry* g) {
    if (V % 4) return false;   // TMA row copies need 16-byte aligned logit rows
    const int P = pairs > 64 ? 4 : (pairs > 32 ? 2 : 1);
    const int R = (pairs + 32 * P - 1) / (32 * P);
    const int H = pairs > 700 ? 4 : 2;
    const int NT = 32 * (R + H);
    if (NT > 1024) return false;
    // (chunk, fetch distance) candidates.  All CTAs should be co-resident (one wave: the
    // kernel is as long as its longest utterance), so the shared-memory budget per CTA is
    // an SM's 227 KB divided by the CTAs per SM the launch needs (at most 4).
    const int tc_env = env().chunk, d_env = env().dist;
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
            g->base = 1;
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
            return pipe_variant(*g) >= 0;
        }
    }
    return false;
}

// Linear-domain kernel: R recursion warps with 8 pairs per thread in registers (256 lattice slots per
// warp), 2 (or 1) combine warps per recursion warp, H softmax / gradient warps.  Needs S_max + 8 <=
// 256 * R slots (alignment shift).  Rows that are not 16-byte aligned (V % 4 != 0: the reference's V = 177,
// params.py:27) use the run-time-stride variants with 4-byte copies.
bool pick_lin(int T, int V, int S_max, int n_utt, Geometry* g) {
    const int P = 8;   // one recursion warp covers 248 labels; also the fastest choice for short targets
    int R = (S_max + P + 32 * P - 1) / (32 * P);
    const bool al = V % 4 == 0;
    // three recursion warps have no instantiation with compile-time strides; the four-warp one (C3's) is faster although
    // its lattice rows are a third wider (V = 48, S ~ 500 ... 760: 356 ns per frame and utterance with three warps and
    // run-time strides, 279 with four; tools/gpu_cliffs.py)
    if (R == 3 && al && V <= 60) R = 4;
    // one helper warp (softmax, then gradient rows) when one recursion warp suffices: 4 warps per CTA
    // leave 128 registers per thread, which the two-rows-in-flight combine pass needs
    // helper warps: one (softmax, then gradient rows) for the narrow vocabularies; two softmax + two gradient warps
    // for rows of more than 64 classes (one helper keeps at most 4 x 64 bit of a frame per lane in registers) and
    // for wide rows that are not 16-byte aligned (the reference's own V = 177, params.py:27: 0.82 ms with one
    // helper and the looped passes, 0.26 ms with four and the MID instantiation on B = 64, T = 750)
    // (rows that are not 16-byte aligned take the four-helper MID passes whatever their width: the one-helper general code
    // copies and normalises them element by element -- V = 29, characters + blank: 0.22 ms against 0.05 ms for V = 48)
    int H = R == 1 ? ((V > 64 || !al) ? 4 : 1) : (V > 256 ? 4 : 2);
    // the MID vocabularies (60 < V <= 256) in a launch that leaves every CTA an SM of its own -- at most 74 utterances: the
    // reference trains with batches of 32 / 64 (deepspeech_ctc/train.py:75-100) at V = 177 -- get EIGHT helper warps,
    // a warp per frame of a chunk: with 7 warps on an SM the helpers' dependent chains bound every iteration
    // (B = 64, T = 750, V = 177: SOFT 2177 / GRAD 2827 busy cycles per chunk against REC 1181 / COMB 1665)
    // ... and so does every other vocabulary of up to 256 classes but the headline one (V = 48 has steady-state loops of
    // its own for all four roles): aligned V <= 64 runs 0.11 ... 0.13 ms there against 0.07 ms (T = 300, B <= 74)
    const bool headline = al && V <= 48 && !env().nofix;   // (V < 48: the same code with a run-time vocabulary, VRUN)
    if (R == 1 && V <= 256 && 2 * std::max(n_utt, 1) <= kNumSmsHint && (H == 4 || !headline)) H = 8;
    if (env().helpers == 1 || env().helpers == 2 || env().helpers == 4 || env().helpers == 8) H = env().helpers;   // developer knob
    const int NC = R <= 4 ? 2 : 1;   // two combine groups while the CTA stays within 512 threads
    const int NT = 32 * ((1 + NC) * R + H + (H == 8 ? 4 : 0));   // (eight helpers come with four copy warps: ctc_lin.cuh, MIDC)
    if (NT > 1024) return false;
    const int NP = 32 * P * R;
    const int RS = lin_row_stride_host(NP, P);
    // fixed emission-ring row stride (an immediate in the kernel): instantiated for 1, 2 and 4 recursion warps
    const int YS = (al && V <= 60 && (R == 1 || R == 2 || R == 4) && !(R == 1 && H >= 4)) ? 80 : 0;
    const int tc_env = env().chunk;
    const int cand[3] = {4, 2, 1};
    for (int pass = 0; pass < 2; ++pass) {
        for (int ci = 0; ci < 3; ++ci) {
            int TC = cand[ci];
            if (tc_env >= 1 && tc_env <= 4) TC = tc_env;   // the kernel unrolls 4 rows
            if (YS == 80 && TC != 4) continue;   // the fixed-stride variants are built for chunks of 4 frames
            // wide aligned rows: a whole warp per frame (chunks of 2 frames with two softmax warps) keeps a row in
            // registers; chunks of 4 would halve the lanes per frame and fall back to the looped passes
            // (nor chunks of 1: the general code is three times slower than a second wave of the WIDE instantiation --
            // V = 2048, T = 300: 0.30 ms at B = 74, 0.97 ms at B = 75)
            if (V > 256 && al && R == 1 && H == 4 && TC != 2 && tc_env == 0) continue;
            const int total = lin_smem_size(NP, R, V, TC, RS, YS);
            // co-resident CTAs per SM a one-wave launch would need -- but never more than the register file holds (every
            // instantiation is built for 128 registers per thread): a 224-thread CTA sits two to an SM whatever its shared
            // memory, so asking it to fit four only pushed the 61 ... 256-class vocabularies off their instantiation (chunks
            // of 2 frames, the general code) from 223 utterances on (V = 177, T = 750: B = 222 0.37 ms, B = 256 1.26 ms)
            const int by_regs = std::max(1, 65536 / (NT * 128));
            const int need = std::max(1, std::min(std::min(4, by_regs), (2 * std::max(n_utt, 1) + kNumSmsHint - 1) / kNumSmsHint));
            const int limit = pass == 0 ? std::min(kMaxSmemBytes, 227 * 1024 / need - 1024) : kMaxSmemBytes;
            if (total > limit) continue;
            g->lP = P; g->lNT = NT; g->lNP = NP; g->lchunk = TC; g->lRS = RS; g->lsmem = total;
            g->lR = R; g->lH = H; g->lD = NC; g->lYS = YS;
            // (at most two co-resident CTAs per SM: 148 utterances on 148 SMs)
            g->lriss = env().rec_iss >= 0 ? env().rec_iss : (std::max(n_utt, 1) <= kNumSmsHint ? 1 : 0);
            g->l_lat_utt_stride = (size_t)std::max(T, 1) * (size_t)RS;
            return lin_variant(*g, V) >= 0;
        }
    }
    return false;
}

int pick_geometry(int T, int V, int S_max, int n_utt, Geometry* g) {
    if (T < 0 || V < 1 || S_max < 0) return CTC_B200_INVALID_ARGUMENT;
    const int pairs = S_max + 1;
    if (pairs > 4096) return CTC_B200_UNSUPPORTED;  // targets longer than 4095 labels
    const char force = env().kernel;   // g: generic, p: log-domain pipe, default: linear
    std::memset(g, 0, sizeof(*g));
    // the log-domain kernel: warp-specialised when the logit rows are 16-byte aligned, else generic
    if (force == 'g' || !pick_pipe(T, V, pairs, n_utt, g)) {
        const int rc = pick_generic(T, V, pairs, g);
        if (rc != CTC_B200_OK) return rc;
    }
    g->pipe = g->base;
    if (force != 'g' && force != 'p' && pick_lin(T, V, S_max, n_utt, g)) g->pipe = 2;
    return CTC_B200_OK;
}

struct Layout { long long frame_stride, utt_stride; };
inline Layout make_layout(int layout, int T, int N, int V) {
    if (layout == CTC_B200_LAYOUT_NTV) return {(long long)V, (long long)T * V};
    return {(long long)N * V, (long long)V};
}

int launch_fused(const float* acts, const int32_t* targets, const int32_t* tgt_offsets,
                 const int32_t* in_lens, const int32_t* tgt_lens, int T, int N, int V, int S_max,
                 int blank, int zero_infinity, int utt_begin, int utt_count, float* nll,
                 float* grad, const float* grad_scale, float* lattice, size_t lattice_bytes,
                 int* status_word, int* flags, int* queue, const ctc_b200_options* opt, cudaStream_t st) {
    if (!acts || !targets || !tgt_offsets || !in_lens || !tgt_lens || !nll || !status_word)
        return CTC_B200_INVALID_ARGUMENT;
    if (T < 0 || N < 0 || V < 1 || blank < 0 || blank >= V || utt_begin < 0 || utt_count < 0 ||
        utt_begin + utt_count > N)
        return CTC_B200_INVALID_ARGUMENT;
    const int layout = opt ? opt->layout : CTC_B200_LAYOUT_TNV;
    if (layout != CTC_B200_LAYOUT_TNV && layout != CTC_B200_LAYOUT_NTV) return CTC_B200_INVALID_ARGUMENT;
    if (opt && opt->use_clamp && !(opt->clamp_min < opt->clamp_max)) return CTC_B200_INVALID_ARGUMENT;
    if (utt_count == 0) return CTC_B200_OK;
    Geometry g;
    int rc = pick_geometry(T, V, S_max, N, &g);
    if (rc != CTC_B200_OK) return rc;
    const size_t lat_need = g.lattice_floats_per_utt() * sizeof(float) * (size_t)utt_count;
    if (!lattice || lattice_bytes < lat_need) return CTC_B200_WORKSPACE_TOO_SMALL;
    if (g.pipe == 2 && !flags) return CTC_B200_INVALID_ARGUMENT;
    // the lattice always, acts / grad whenever the vocabulary allows 128-bit row accesses
    if ((reinterpret_cast<uintptr_t>(lattice) & 15) ||
        (V % 4 == 0 && ((reinterpret_cast<uintptr_t>(acts) & 15) || (grad && (reinterpret_cast<uintptr_t>(grad) & 15)))) ||
        (reinterpret_cast<uintptr_t>(acts) & 3) || (grad && (reinterpret_cast<uintptr_t>(grad) & 3)))
        return CTC_B200_INVALID_ARGUMENT;

    const Layout lay = make_layout(layout, T, N, V);
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
    prm.ysave = nullptr;   // (saved softmax rows: tried, no faster; the kernel keeps the hook)
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
    prm.frame_stride = lay.frame_stride;
    prm.utt_stride = lay.utt_stride;
    prm.use_clamp = (opt && opt->use_clamp) ? 1 : 0;
    prm.clamp_lo = prm.use_clamp ? opt->clamp_min : 0.f;
    prm.clamp_hi = prm.use_clamp ? opt->clamp_max : 0.f;
    prm.redo = nullptr;
    int* fl = flags ? flags - 2 * (ptrdiff_t)utt_begin : nullptr;   // kernels index flags by absolute utterance
    // Forward-only calls (no gradient: torch.no_grad(), validation loss) go straight to the log-domain
    // kernel: the linear kernel's safety net is the per-frame posterior-mass check of its GRADIENT pass,
    // so without a gradient a partial loss of mass (flushed cells under saturated logits) would go unseen.
    const bool use_lin = g.pipe == 2 && grad != nullptr;
    if (use_lin) {
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
        lp.rotate = (env().rotate && (g.lNT / 32) % 4 == 0) ? num_sms() : 0;
        lp.redo = nullptr;
        lp.queue = nullptr;
        lp.n_utt = utt_count;
        lp.utt_rot = 0;
        lp.map_mode = 0;
        lp.map_pairs = std::max(1, num_sms() / 2);
        int n_clusters = utt_count;
        // Optional (opt->persistent): with more utterances than co-resident clusters, launch only those and
        // let them pull utterances from a device-side queue, longest first.  Off by default -- the hardware
        // CTA scheduler already hands the next cluster of the grid to the first free slot, in the same order.
        const bool want_persist = env().persist > 0 || (env().persist < 0 && opt && opt->persistent);
        const int resident = (want_persist && queue && lin_supports_queue(g, V)) ? lin_resident_clusters(g, V) : 0;
        if (resident > 0 && utt_count > resident) {
            lp.queue = queue;
            n_clusters = std::min(resident, utt_count);
        } else {
            // CTA placement for a launch that fits one wave with an incomplete last "layer" (C2: 256 clusters on
            // 74 SM pairs = 3 full layers + 34 clusters): the batch is sorted by length (dataloader.py:53), so with
            // the identity mapping the LONGEST utterances share their SMs with a fourth CTA.  Rotating the
            // utterances by whole layers moves them to the SM pairs that host only three (measured on B200, C2:
            // 0.263 ms -> 0.248 ms with a rotation of two layers; rotations that are not whole layers: 0.27 ms).
            const int pairs = std::max(1, num_sms() / 2), layers = utt_count / pairs;
            int rot = (utt_count > pairs && utt_count <= 4 * pairs && utt_count % pairs != 0)
                          ? pairs * std::max(1, layers - 1) : 0;
            if (env().utt_rot >= 0) rot = env().utt_rot;
            lp.utt_rot = std::max(0, std::min(rot, utt_count - 1));
            if (env().map_mode > 0 && utt_count > pairs && utt_count <= 4 * pairs) lp.map_mode = env().map_mode;
        }
        CTC_CUDA(launch_lin(lp, fl, g, n_clusters, st));
#ifdef CTC_B200_DEV_KNOBS   // developer builds only: look at the linear kernel's own output for flagged utterances
        if (Env::geti("CTC_B200_NOFALLBACK", 0)) return CTC_B200_OK;
#endif
    }
    if (g.base == 1) {
        PipeParams pp;
        pp.f = prm;
        pp.redo = use_lin ? fl : nullptr;
        pp.utt_rot = 0;
        pp.map_mode = 0;
        pp.map_pairs = 0;
        pp.queue = nullptr;
        pp.n_utt = utt_count;
        pp.R = g.R;
        pp.H = g.G;
        pp.NP = g.NP;
        pp.D = g.D;
        pp.rotate = env().rotate ? num_sms() : 0;
        CTC_CUDA(launch_pipe(pp, g, utt_count, pp.redo != nullptr && t_allow_pdl && env().pdl, st));
        return CTC_B200_OK;
    }
    prm.redo = use_lin ? fl : nullptr;
    CTC_CUDA(launch_generic(prm, g, utt_count, st));
    return CTC_B200_OK;
}

int status_from_bits(int bits) {
    if (bits & kStatusBadLabel) return CTC_B200_BAD_LABEL;
    if (bits & kStatusBadLength) return CTC_B200_BAD_LENGTH;
    if (bits & kStatusPeerTimeout) return CTC_B200_PEER_TIMEOUT;
    return CTC_B200_OK;
}

}  // namespace

extern "C" {

int ctc_b200_version(void) { return 2000; }

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
    // [256 B header: status word, utterance queue][redo flags: 2 ints per utterance][lattice]
    out->workspace_bytes = kHeaderBytes + flag_bytes(n_utt) +
                           (g.lattice_floats_per_utt() * sizeof(float) * (size_t)n_utt + 255) / 256 * 256;
    out->variant = lin ? lin_variant(g, V) : (g.pipe == 1 ? pipe_variant(g) : generic_variant(g));
    out->fallback_kernel = lin ? g.base : -1;
    out->comb_groups = lin ? g.lD : 0;
    out->resident_clusters = 0;
    if (lin) {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) out->resident_clusters = lin_resident_clusters(g, V);
        else cudaGetLastError();
    }
    return CTC_B200_OK;
}

const char* ctc_b200_variant_name(int kernel, int variant) {
    switch (kernel) {
        case 2: return lin_variant_name(variant);
        case 1: return pipe_variant_name(variant);
        case 0: return generic_variant_name(variant);
    }
    return "?";
}

int ctc_b200_workspace_bytes(int T, int N, int V, int S_max, size_t* bytes) {
    if (!bytes) return CTC_B200_INVALID_ARGUMENT;
    ctc_b200_geometry g;
    int rc = ctc_b200_get_geometry(T, N, V, S_max, &g);
    if (rc == CTC_B200_OK) *bytes = g.workspace_bytes;
    return rc;
}

int ctc_b200_fwd_bwd_ex_f32(const float* acts, const int32_t* targets, const int32_t* tgt_offsets,
                            const int32_t* in_lens, const int32_t* tgt_lens, int T, int N, int V,
                            int S_max, int blank, int zero_infinity, int utt_begin, int utt_count,
                            float* nll, float* grad, const float* grad_scale, void* workspace,
                            size_t workspace_bytes, const ctc_b200_options* opt, void* stream) {
    const size_t head = (size_t)kHeaderBytes + flag_bytes(utt_count);
    if (!workspace || workspace_bytes < head) return CTC_B200_WORKSPACE_TOO_SMALL;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return CTC_B200_INVALID_ARGUMENT;
    char* ws = static_cast<char*>(workspace);
    return launch_fused(acts, targets, tgt_offsets, in_lens, tgt_lens, T, N, V, S_max, blank,
                        zero_infinity, utt_begin, utt_count, nll, grad, grad_scale,
                        reinterpret_cast<float*>(ws + head), workspace_bytes - head,
                        reinterpret_cast<int*>(ws), reinterpret_cast<int*>(ws + kHeaderBytes),
                        reinterpret_cast<int*>(ws + kQueueOffset), opt, static_cast<cudaStream_t>(stream));
}

int ctc_b200_fwd_bwd_range_f32(const float* acts, const int32_t* targets,
                               const int32_t* tgt_offsets, const int32_t* in_lens,
                               const int32_t* tgt_lens, int T, int N, int V, int S_max,
                               int blank, int zero_infinity, int utt_begin, int utt_count,
                               float* nll, float* grad, const float* grad_scale,
                               void* workspace, size_t workspace_bytes, void* stream) {
    return ctc_b200_fwd_bwd_ex_f32(acts, targets, tgt_offsets, in_lens, tgt_lens, T, N, V, S_max, blank,
                                   zero_infinity, utt_begin, utt_count, nll, grad, grad_scale, workspace,
                                   workspace_bytes, nullptr, stream);
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

int ctc_b200_scale_grad_ex_f32(float* grad, const float* scale, int per_utt, int T, int N, int V,
                               int layout, void* stream) {
    if (!grad || !scale || T < 0 || N < 0 || V < 1) return CTC_B200_INVALID_ARGUMENT;
    if (layout != CTC_B200_LAYOUT_TNV && layout != CTC_B200_LAYOUT_NTV) return CTC_B200_INVALID_ARGUMENT;
    if (T == 0 || N == 0) return CTC_B200_OK;
    const int threads = 256;
    const size_t per_utt_elems = (size_t)T * V;
    int gy = (int)std::min<size_t>((per_utt_elems + threads * 8 - 1) / (threads * 8), 64);
    gy = std::max(gy, 1);
    const Layout lay = make_layout(layout, T, N, V);
    ctc_scale_grad_kernel<<<dim3(N, gy), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        grad, scale, per_utt, T, N, V, lay.frame_stride, lay.utt_stride);
    CTC_CUDA(cudaGetLastError());
    return CTC_B200_OK;
}

int ctc_b200_scale_grad_f32(float* grad, const float* scale, int per_utt, int T, int N, int V,
                            void* stream) {
    return ctc_b200_scale_grad_ex_f32(grad, scale, per_utt, T, N, V, CTC_B200_LAYOUT_TNV, stream);
}

static int launch_reduce(const float* nll, const int32_t* in_lens, const int32_t* tgt_lens, int N,
                         int reduction, int zero_on_short, float* out2, float* loss, float* result4,
                         void* stream) {
    if (!nll || !tgt_lens || !out2 || N < 0) return CTC_B200_INVALID_ARGUMENT;
    const int mode = reduction == CTC_B200_REDUCE_MEAN ? 1 : 2;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (t_allow_pdl && env().pdl)
        CTC_CUDA(launch_pdl(ctc_reduce_loss_kernel, dim3(1), dim3(256), (size_t)0, st, nll, in_lens, tgt_lens, N,
                            mode, zero_on_short, out2, loss, result4));
    else
        ctc_reduce_loss_kernel<<<1, 256, 0, st>>>(nll, in_lens, tgt_lens, N, mode, zero_on_short, out2, loss, result4);
    CTC_CUDA(cudaGetLastError());
    return CTC_B200_OK;
}

int ctc_b200_reduce_loss_f32(const float* nll, const int32_t* tgt_lens, int N, int reduction,
                             float* out2, float* loss, void* stream) {
    return launch_reduce(nll, nullptr, tgt_lens, N, reduction, 0, out2, loss, nullptr, stream);
}

int ctc_b200_reduce_loss_status_f32(const float* nll, const int32_t* in_lens, const int32_t* tgt_lens,
                                    int N, int reduction, int zero_on_short, float* out2,
                                    float* loss, float* result4, void* stream) {
    if (!result4) return CTC_B200_INVALID_ARGUMENT;
    return launch_reduce(nll, in_lens, tgt_lens, N, reduction, zero_on_short, out2, loss, result4, stream);
}

static std::atomic<long long> g_peer_timeout_ms{600000};

long long ctc_b200_set_peer_timeout_ms(long long ms) {
    return g_peer_timeout_ms.exchange(ms < 0 ? 0 : ms);
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
    const unsigned long long timeout_ns = (unsigned long long)g_peer_timeout_ms.load() * 1000000ull;
    const int threads = pair_in ? 32 : 256;   // the exchange-only form needs one warp
    if (t_allow_pdl && env().pdl)
        CTC_CUDA(launch_pdl(ctc_reduce_loss_allreduce_kernel, dim3(1), dim3(threads), (size_t)0, st, nll, tgt_lens,
                            N, mode, pair_in, pb, rank, world_size, seq, timeout_ns, out2, loss,
                            static_cast<int*>(workspace)));
    else
        ctc_reduce_loss_allreduce_kernel<<<1, threads, 0, st>>>(nll, tgt_lens, N, mode, pair_in, pb, rank, world_size,
                                                                seq, timeout_ns, out2, loss, static_cast<int*>(workspace));
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

int ctc_b200_greedy_decode_ler_i32(const float* acts, int T, int N, int V, int layout,
                                   const int32_t* in_lens, const int32_t* targets,
                                   const int32_t* tgt_offsets, const int32_t* tgt_lens, int blank,
                                   int32_t* hyp, int32_t* hyp_len, int32_t* dist,
                                   long long* totals, void* stream) {
    if (!acts || !in_lens || !hyp || !hyp_len || T < 0 || N < 0 || V < 1 || blank < 0 || blank >= V)
        return CTC_B200_INVALID_ARGUMENT;
    if (layout != CTC_B200_LAYOUT_TNV && layout != CTC_B200_LAYOUT_NTV) return CTC_B200_INVALID_ARGUMENT;
    if (targets && (!tgt_offsets || !tgt_lens || !dist)) return CTC_B200_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(acts) & 3) return CTC_B200_INVALID_ARGUMENT;
    const Layout lay = make_layout(layout, T, N, V);
    CTC_CUDA(launch_decode_ler(acts, T, N, V, lay.frame_stride, lay.utt_stride, in_lens, targets, tgt_offsets,
                               tgt_lens, blank, hyp, hyp_len, dist, totals, static_cast<cudaStream_t>(stream)));
    return CTC_B200_OK;
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
    int* d_queue = nullptr;    // [2 * n_slices] utterance queues (persistent launches), zeroed once
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
    long long last_h2d_bytes = 0;
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
    ok(cudaMalloc(&s->d_queue, (size_t)s->n_slices * 2 * sizeof(int)));
    if (e == cudaSuccess) ok(cudaMemset(s->d_queue, 0, (size_t)s->n_slices * 2 * sizeof(int)));
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
    cudaFree(s->d_acts); cudaFree(s->d_grad); cudaFree(s->d_lattice); cudaFree(s->d_flags); cudaFree(s->d_queue);
    cudaFree(s->d_small); cudaFree(s->d_res);
    if (s->h_small) cudaFreeHost(s->h_small);
    if (s->h_res) cudaFreeHost(s->h_res);
    delete s;
    return CTC_B200_OK;
}

float* ctc_b200_session_grad_device(ctc_b200_session* s) { return s ? s->d_grad : nullptr; }
int ctc_b200_session_last_launches(const ctc_b200_session* s) { return s ? s->last_launches : 0; }
long long ctc_b200_session_last_h2d_bytes(const ctc_b200_session* s) { return s ? s->last_h2d_bytes : 0; }

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
    long long h2d = (long long)s->small_bytes;
    const size_t pitch = (size_t)N * V * sizeof(float);
    for (int k = 0; k < s->n_slices; ++k) {
        const int b0 = (int)((long long)N * k / s->n_slices);
        const int b1 = (int)((long long)N * (k + 1) / s->n_slices);
        if (b1 == b0) continue;
        // variable-length masking needs no padding: frames t >= T_b are never read, so a slice's copy stops at
        // its longest utterance (the reference's batches are sorted by length, dataloader.py:53: the slices of
        // short utterances move correspondingly fewer rows)
        int rows = 0;
        for (int b = b0; b < b1; ++b) rows = std::max(rows, in_lens_host[b]);
        if (rows > 0)
            CTC_CUDA(cudaMemcpy2DAsync(s->d_acts + (size_t)b0 * V, pitch, acts_host + (size_t)b0 * V,
                                       pitch, (size_t)(b1 - b0) * V * sizeof(float), (size_t)rows,
                                       cudaMemcpyHostToDevice, s->s_copy));
        h2d += (long long)rows * (b1 - b0) * V * (long long)sizeof(float);
        CTC_CUDA(cudaEventRecord(s->ev[k], s->s_copy));
        cudaStream_t sk = env().slice_streams ? s->s_slice[k] : s->s_comp;
        CTC_CUDA(cudaStreamWaitEvent(sk, s->ev[s->n_slices], 0));   // targets / lengths / cleared status
        CTC_CUDA(cudaStreamWaitEvent(sk, s->ev[k], 0));              // this slice's logits
        t_allow_pdl = false;
        int rc = launch_fused(s->d_acts, d_tg, d_off, d_il, d_tl, T, N, V, s->S_max, blank,
                              zero_infinity, b0, b1 - b0, d_nll, want_grad ? s->d_grad : nullptr,
                              d_sc, s->d_lattice + (size_t)b0 * s->geo.lattice_floats_per_utt(),
                              s->lattice_bytes - (size_t)b0 * s->geo.lattice_floats_per_utt() * sizeof(float),
                              d_status, s->d_flags + 2 * (size_t)b0, s->d_queue + 2 * k, nullptr, sk);
        t_allow_pdl = true;
        if (rc != CTC_B200_OK) return rc;
        CTC_CUDA(cudaEventRecord(s->ev_done[k], sk));
        CTC_CUDA(cudaStreamWaitEvent(s->s_comp, s->ev_done[k], 0));
        launches += (s->geo.pipe == 2 && want_grad) ? 2 : 1;
    }
    ctc_reduce_loss_kernel<<<1, 256, 0, s->s_comp>>>(
        d_nll, nullptr, d_tl, N, reduction == CTC_B200_REDUCE_MEAN ? 1 : 2, 0, d_out2, nullptr, nullptr);
    CTC_CUDA(cudaGetLastError());
    ++launches;
    CTC_CUDA(cudaMemcpyAsync(s->h_res, s->d_res, nll_host ? s->res_bytes : 16,
                             cudaMemcpyDeviceToHost, s->s_comp));
    if (grad_host && want_grad)
        CTC_CUDA(cudaMemcpyAsync(grad_host, s->d_grad, (size_t)T * N * V * sizeof(float),
                                 cudaMemcpyDeviceToHost, s->s_comp));
    CTC_CUDA(cudaStreamSynchronize(s->s_comp));
    s->last_launches = launches;
    s->last_h2d_bytes = h2d;

    const float* h_out2 = reinterpret_cast<const float*>(s->h_res);
    const int bits = *reinterpret_cast<const int*>(s->h_res + 8);
    *loss_host = (reduction == CTC_B200_REDUCE_MEAN) ? h_out2[0] / (float)N : h_out2[0];
    if (nll_host) std::memcpy(nll_host, s->h_res + 16, (size_t)N * 4);
    return status_from_bits(bits);
}

}  // extern "C"
