// ctc_decode.cu -- greedy CTC decode + label-error count on the GPU.
//
// Replaces, for `validate`, the host-side Python loops of the reference:
//   asr/models/trainer.py:450-463  unit_validate: onehot2int (arg-max over the vocabulary,
//                                  asr/utils/misc.py:44-51) of every frame t < frame_lens[b], then
//                                  remove_duplicates(blank=0) (misc.py:78-84: collapse repeats, drop blanks)
//   asr/models/trainer.py:336-343  edit_distance: Levenshtein distance (unit costs) per utterance, summed
//   asr/models/trainer.py:301-304  LER = 100 * sum(distance) / sum(len(ref))
//
// Two kernels:
//   ctc_argmax_kernel        HBM-bound: every valid row of V logits is read once (a warp per row, 128-bit
//                            loads when rows are 16-byte aligned), 4 bytes out per row.  Algorithmic bytes:
//                            4 * V * sum_b T_b read + 4 * sum_b T_b written.
//   ctc_collapse_ler_kernel  one CTA per utterance: in-place stream compaction of the arg-max row
//                            (ballot / popc scan), then the Levenshtein distance.  References of up to 1024 labels
//                            whose match masks (V x ceil(S / 32) words) fit the shared-memory budget: Myers' bit-vector
//                            algorithm (J. ACM 46(3), 1999; the block form with horizontal carries), one 32-bit block
//                            of the reference per lane of ONE warp, the blocks of a hypothesis label processed
//                            systolically (lane k works on label t - k in step t and takes the carry lane k - 1
//                            produced one step earlier): H + S / 32 steps of ~25 integer instructions, no barrier.
//                            Otherwise: anti-diagonals (one CTA barrier per diagonal, three diagonals of S+1 ints).
#include "ctc_launch.h"

#include <climits>

namespace ctcb200 {

namespace {

constexpr int kDecodeThreads = 256;
constexpr int kMaxRef = 4095;   // as the loss: targets longer than 4095 labels are unsupported

// rows are enumerated in MEMORY order (time-major: t * N + b; batch-major: b * T + t), a warp per row
__global__ void ctc_argmax_kernel(const float* __restrict__ acts, int T, int N, int V,
                                  long long frame_stride, long long utt_stride, int bmajor,
                                  const int32_t* __restrict__ in_lens, int32_t* __restrict__ best,
                                  long long* __restrict__ totals) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    if (blockIdx.x == 0 && threadIdx.x < 2 && totals != nullptr) totals[threadIdx.x] = 0;
    const bool al = (V & 3) == 0 && ((frame_stride | utt_stride) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(acts) & 15) == 0;
    const long long rows = (long long)T * N;
    // narrow rows: a group of G = 2^k lanes per row (G x 128 bit >= the row, or G floats for rows that are not 16-byte
    // aligned), 32 / G rows per warp -- a warp per row would leave 20 of 32 lanes idle at V = 48
    const int per_lane_units = al ? (V >> 2) : V;
    int G = 32;
    while (G > 1 && (G >> 1) >= per_lane_units) G >>= 1;
    const int rpw = 32 / G, gl = lane & (G - 1), sub = lane / G;
    for (long long r0 = warp0 * rpw; r0 < rows; r0 += nwarps * rpw) {
        const long long r = r0 + sub;
        bool live = r < rows;
        int t = 0, b = 0;
        if (live) {
            t = bmajor ? (int)(r % T) : (int)(r / N);
            b = bmajor ? (int)(r / T) : (int)(r % N);
            live = t < in_lens[b];
        }
        float v = -CUDART_INF_F;
        int idx = INT_MAX;
        if (live) {
            const float* row = acts + (size_t)t * (size_t)frame_stride + (size_t)b * (size_t)utt_stride;
            if (al) {
                const float4* row4 = reinterpret_cast<const float4*>(row);
                for (int c = gl; c < (V >> 2); c += G) {
                    const float4 x = __ldg(row4 + c);
                    // ascending index inside the lane, strict '>': the first maximum is kept
                    if (x.x > v || idx == INT_MAX) { v = x.x; idx = 4 * c; }
                    if (x.y > v) { v = x.y; idx = 4 * c + 1; }
                    if (x.z > v) { v = x.z; idx = 4 * c + 2; }
                    if (x.w > v) { v = x.w; idx = 4 * c + 3; }
                }
            } else {
                for (int c = gl; c < V; c += G) {
                    const float x = __ldg(row + c);
                    if (x > v || idx == INT_MAX) { v = x; idx = c; }
                }
            }
        }
        for (int o = G >> 1; o > 0; o >>= 1) {      // (xor partners stay inside the group: G is a power of two)
            const float vo = __shfl_xor_sync(0xffffffffu, v, o);
            const int io = __shfl_xor_sync(0xffffffffu, idx, o);
            if (io != INT_MAX && (idx == INT_MAX || vo > v || (vo == v && io < idx))) { v = vo; idx = io; }
        }
        if (live && gl == 0) best[(size_t)b * T + t] = idx == INT_MAX ? 0 : idx;
    }
}

__global__ void __launch_bounds__(kDecodeThreads)
ctc_collapse_ler_kernel(int T, const int32_t* __restrict__ in_lens, const int32_t* __restrict__ targets,
                        const int32_t* __restrict__ tgt_off, const int32_t* __restrict__ tgt_lens,
                        int blank, int32_t* __restrict__ hyp, int32_t* __restrict__ hyp_len,
                        int32_t* __restrict__ dist, long long* __restrict__ totals, int stage_ints, int V) {
    extern __shared__ __align__(16) int s_diag[];   // 3 x (S + 1) ints [+ hypothesis + reference]: stage_ints in all
    __shared__ int s_warp[kDecodeThreads / 32];
    __shared__ int s_last, s_out;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int Tb = min(max(in_lens[b], 0), T);
    int32_t* row = hyp + (size_t)b * T;

    // ---- collapse repeats, drop blanks: in-place compaction, a tile of blockDim.x frames per pass ----
    if (tid == 0) { s_last = -1; s_out = 0; }
    __syncthreads();
    for (int base = 0; base < Tb; base += kDecodeThreads) {
        const int t = base + tid;
        const int cur = t < Tb ? row[t] : blank;
        int prev = __shfl_up_sync(0xffffffffu, cur, 1);
        if (lane == 0) prev = (tid == 0) ? s_last : (t - 1 < Tb ? row[t - 1] : blank);
        const bool keep = t < Tb && cur != blank && cur != prev;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[w] = __popc(bal);
        __syncthreads();                      // every read of this tile is done; warp counts are posted
        int off = s_out;
        for (int i = 0; i < w; ++i) off += s_warp[i];
        off += __popc(bal & ((1u << lane) - 1u));
        if (keep) row[off] = cur;             // off <= t: never ahead of a frame that is still to be read
        __syncthreads();
        if (tid == kDecodeThreads - 1) {
            s_last = cur;                     // (only read when another tile follows: then t < Tb)
            int tot = s_out;
            for (int i = 0; i < kDecodeThreads / 32; ++i) tot += s_warp[i];
            s_out = tot;
        }
        __syncthreads();
    }
    const int H = s_out;
    if (tid == 0) hyp_len[b] = H;
    if (targets == nullptr) return;

    // ---- Levenshtein distance (unit costs) by anti-diagonals ---------------------------------------
    // D[i][j]: hyp[:i] vs ref[:j].  Diagonal d holds the cells with i + j = d at index j.
    const int S = tgt_lens[b];
    if (S < 0 || S > kMaxRef) {
        if (tid == 0) dist[b] = -1;
        return;
    }
    const int32_t* ref = targets + tgt_off[b];
    // ---- Myers' bit-vector algorithm: block k (reference labels 32 k ... 32 k + 31) lives in lane k of warp 0 ----------
    const int W = (S + 31) >> 5;
    if (S >= 1 && W <= 32 && (long long)V * W + H <= (long long)stage_ints) {
        unsigned* peq = reinterpret_cast<unsigned*>(s_diag);       // [V][W]: bit j of peq[c][k] = (ref[32 k + j] == c)
        int* s_h = s_diag + V * W;                                 // the hypothesis
        for (int i = tid; i < V * W; i += kDecodeThreads) peq[i] = 0u;
        for (int i = tid; i < H; i += kDecodeThreads) s_h[i] = row[i];
        __syncthreads();
        for (int j = tid; j < S; j += kDecodeThreads) {
            const int c = ref[j];
            if (c >= 0 && c < V) atomicOr(&peq[c * W + (j >> 5)], 1u << (j & 31));   // (a label outside [0, V) matches nothing)
        }
        __syncthreads();
        if (w != 0) return;
        const int k = lane;                        // my block
        const bool mine = k < W;
        const unsigned top = k == W - 1 ? 1u << ((S - 1) & 31) : 0x80000000u;
        unsigned Pv = 0xffffffffu, Mv = 0u;
        int score = S, hout = 0;
        // branch-free steps (selects only: the lanes are at different labels, and at the ends of the sweep some have
        // none); the match mask of the NEXT step's label is loaded one step ahead, off the carry chain
        const int kk = mine ? k : 0;
        auto mask_of = [&](int i) -> unsigned {
            const int ic = min(max(i, 0), max(H - 1, 0));
            const int c = H > 0 ? s_h[ic] : -1;
            return (c >= 0 && c < V) ? peq[c * W + kk] : 0u;
        };
        unsigned Eq_next = mask_of(-k);
        for (int t = 0; t < (H > 0 ? H + W - 1 : 0); ++t) {
            // the horizontal carry of hypothesis label t - k: what lane k - 1 produced for it one step ago; the first
            // block takes +1 (D[i][0] = i)
            int hin = __shfl_up_sync(0xffffffffu, hout, 1);
            if (k == 0) hin = 1;
            const bool active = mine && (unsigned)(t - k) < (unsigned)H;
            unsigned Eq = Eq_next;
            Eq_next = mask_of(t + 1 - k);
            const unsigned neg = hin < 0 ? 1u : 0u, pos = hin > 0 ? 1u : 0u;
            const unsigned Xv = Eq | Mv;
            Eq |= neg;
            const unsigned Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            unsigned Ph = Mv | ~(Xh | Pv);
            unsigned Mh = Pv & Xh;
            const int h = (Ph & top) ? 1 : ((Mh & top) ? -1 : 0);
            Ph = (Ph << 1) | pos;
            Mh = (Mh << 1) | neg;
            const unsigned Pn = Mh | ~(Xv | Ph), Mn = Ph & Xv;
            Pv = active ? Pn : Pv;
            Mv = active ? Mn : Mv;
            hout = active ? h : 0;
            score += (k == W - 1) ? hout : 0;
        }
        const int dval = __shfl_sync(0xffffffffu, score, W - 1);
        if (lane == 0) {
            dist[b] = dval;
            if (totals != nullptr) {
                atomicAdd(reinterpret_cast<unsigned long long*>(totals), (unsigned long long)dval);
                atomicAdd(reinterpret_cast<unsigned long long*>(totals) + 1, (unsigned long long)S);
            }
        }
        return;
    }
    int* p2 = s_diag;                 // diagonal d - 2
    int* p1 = s_diag + (S + 1);       // diagonal d - 1
    int* cu = s_diag + 2 * (S + 1);   // diagonal d
    // hypothesis and reference are read once per cell: staged in shared memory when they fit behind the diagonals
    // (a global load per cell and diagonal was most of the 0.27 us a diagonal took)
    const int32_t* hp = row;
    const int32_t* rp = ref;
    if (3 * (S + 1) + H + S <= stage_ints) {
        int* s_h = s_diag + 3 * (S + 1);
        int* s_r = s_h + H;
        for (int i = tid; i < H; i += kDecodeThreads) s_h[i] = row[i];
        for (int j = tid; j < S; j += kDecodeThreads) s_r[j] = ref[j];
        hp = s_h;
        rp = s_r;
        __syncthreads();
    }
    for (int d = 0; d <= H + S; ++d) {
        for (int j = tid; j <= S; j += kDecodeThreads) {
            const int i = d - j;
            if (i < 0 || i > H) continue;
            int v;
            if (i == 0) v = j;
            else if (j == 0) v = i;
            else {
                const int sub = p2[j - 1] + (hp[i - 1] != rp[j - 1] ? 1 : 0);
                v = min(min(p1[j] + 1, p1[j - 1] + 1), sub);
            }
            cu[j] = v;
        }
        __syncthreads();
        int* tmp = p2; p2 = p1; p1 = cu; cu = tmp;
    }
    if (tid == 0) {
        const int dval = p1[S];       // the last diagonal written (d = H + S) holds D[H][S] at index S
        dist[b] = dval;
        if (totals != nullptr) {
            atomicAdd(reinterpret_cast<unsigned long long*>(totals), (unsigned long long)dval);
            atomicAdd(reinterpret_cast<unsigned long long*>(totals) + 1, (unsigned long long)S);
        }
    }
}

}  // namespace

cudaError_t launch_decode_ler(const float* acts, int T, int N, int V, long long frame_stride,
                              long long utt_stride, const int32_t* in_lens, const int32_t* targets,
                              const int32_t* tgt_off, const int32_t* tgt_lens, int blank, int32_t* hyp,
                              int32_t* hyp_len, int32_t* dist, long long* totals, cudaStream_t st) {
    if (T == 0 || N == 0) {
        if (totals) return last_cuda_error_set(cudaMemsetAsync(totals, 0, 2 * sizeof(long long), st));
        return cudaSuccess;
    }
    const long long rows = (long long)T * N;
    // 8 rows per CTA of 256 threads; enough CTAs to fill the machine a few times over, grid-stride beyond
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)num_sms() * 32;
    if (blocks > cap) blocks = cap;
    const int bmajor = frame_stride < utt_stride ? 1 : 0;
    ctc_argmax_kernel<<<dim3((unsigned)blocks), dim3(256), 0, st>>>(acts, T, N, V, frame_stride, utt_stride,
                                                                  bmajor, in_lens, hyp, totals);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return last_cuda_error_set(e);
    // three diagonals of up to kMaxRef + 1 ints (48 KB of dynamic shared memory on top of the static part), and room for
    // the hypothesis (at most T labels) and the reference behind them where that stays within 160 KB
    const size_t diag = 3 * (size_t)(kMaxRef + 1) * sizeof(int);
    const size_t smem = std::min<size_t>(160 * 1024, diag + ((size_t)T + kMaxRef) * sizeof(int));
    static SmemMark mark;
    e = ensure_smem(reinterpret_cast<const void*>(ctc_collapse_ler_kernel), mark, (int)smem);
    if (e != cudaSuccess) return e;
    ctc_collapse_ler_kernel<<<dim3(N), dim3(kDecodeThreads), smem, st>>>(T, in_lens, targets, tgt_off, tgt_lens,
                                                                        blank, hyp, hyp_len, dist, totals,
                                                                        (int)(smem / sizeof(int)), V);
    return last_cuda_error_set(cudaGetLastError());
}

}  // namespace ctcb200
