// ctc_small.cuh -- the small kernels around the fused kernel: gradient scaling (autograd's grad_output),
// loss reduction (+ the trainer's post-loss host checks), and the loss reduction fused with the
// data-parallel job's only collective.  Included by ctc_abi.cu only.
#pragma once
#include "ctc_kernels.cuh"

namespace ctcb200 {

// ---------------------------------------------------------------------------
// grad[t, b, :] *= scale[b] (or *= scale[0] when `per_utt` is 0), skipped
// entirely -- no memory traffic -- for factors equal to 1.  This is the whole
// "backward": the fused kernel already wrote d nll / d logits; autograd only
// has to apply the upstream grad_output (trainer.py:429 loss.mul_(0), AMP loss
// scaling at :435-436), which is exactly 1 in the plain fp32 step.
// Element (t, b, v) lives at t * frame_stride + b * utt_stride + v.
// ---------------------------------------------------------------------------
__global__ void ctc_scale_grad_kernel(float* __restrict__ grad, const float* __restrict__ scale,
                                      int per_utt, int T, int N, int V, long long frame_stride,
                                      long long utt_stride) {
    const int b = blockIdx.x;
    const float s = per_utt ? scale[b] : scale[0];
    if (s == 1.0f) return;
    const size_t total = (size_t)T * V;
    float* gb = grad + (size_t)b * (size_t)utt_stride;
    for (size_t idx = (size_t)blockIdx.y * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.y * blockDim.x) {
        const size_t t = idx / V, v = idx - t * V;
        float* g = gb + t * (size_t)frame_stride + v;
        *g = (s == 0.0f) ? 0.0f : *g * s;
    }
}

// ---------------------------------------------------------------------------
// out[0] = sum_b nll_b / max(S_b, 1)   (mode 1, 'mean' numerator; trainer.py:153)
//        = sum_b nll_b                 (mode 2, 'sum')
// out[1] = N  (the normaliser that is all-reduced together with out[0])
// Single CTA, fixed summation order => bit-reproducible.
//
// result4 != nullptr: also the trainer's post-loss host checks (trainer.py:423-430), so that one
// 16-byte read replaces three device->host syncs: [loss, flags (int bits), backward factor, #short].
// ---------------------------------------------------------------------------
__global__ void ctc_reduce_loss_kernel(const float* __restrict__ nll, const int32_t* __restrict__ in_lens,
                                       const int32_t* __restrict__ tgt_lens, int N, int mode,
                                       int zero_on_short, float* __restrict__ out,
                                       float* __restrict__ loss, float* __restrict__ result4) {
    __shared__ double s_part[32];
    __shared__ int s_short[32];
    // may be launched with programmatic stream serialization: the producer of `nll` must be complete
    asm volatile("griddepcontrol.wait;" ::: "memory");
    double acc = 0.0;
    int n_short = 0;
    for (int b = threadIdx.x; b < N; b += blockDim.x) {
        double v = (double)nll[b];
        const int S = tgt_lens[b];
        if (mode == 1) v /= (double)max(S, 1);
        acc += v;
        if (in_lens != nullptr && in_lens[b] < 2 * S) ++n_short;   // trainer.py:427 frame_lens < 2 * label_lens
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        n_short += __shfl_xor_sync(0xffffffffu, n_short, o);
    }
    if ((threadIdx.x & 31) == 0) { s_part[threadIdx.x >> 5] = acc; s_short[threadIdx.x >> 5] = n_short; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        int ns = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { s += s_part[i]; ns += s_short[i]; }
        out[0] = (float)s;
        out[1] = (float)N;
        const float l = (mode == 1) ? (float)(s / (double)max(N, 1)) : (float)s;
        if (loss && !result4) loss[0] = l;
        if (result4) {
            int flags = 0;
            if (l != l) flags |= CTC_B200_FLAG_NAN;
            if (l == CUDART_INF_F || l == -CUDART_INF_F) flags |= CTC_B200_FLAG_INF;
            if (ns > 0) flags |= CTC_B200_FLAG_SHORT;
            const bool zero = zero_on_short && ns > 0 && !(flags & (CTC_B200_FLAG_NAN | CTC_B200_FLAG_INF));
            result4[0] = zero ? 0.0f : l;
            if (loss) loss[0] = zero ? 0.0f : l;
            result4[1] = (float)flags;
            result4[2] = zero ? 0.0f : 1.0f;
            result4[3] = (float)ns;
        }
    }
}

// ---------------------------------------------------------------------------
// Loss reduction fused with the job's only collective: the (sum, count) pair goes straight into
// every peer's exchange buffer (P2P stores over NVLink), the pairs addressed to this rank are
// awaited and added in rank order.  Exchange buffer: slot[parity][rank] = {sum, seq, count, seq}.
// ---------------------------------------------------------------------------
struct PeerBufs { float4* p[8]; };

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// pair_in != nullptr: the local pair was already reduced (ctc_reduce_loss_kernel); exchange only.
// timeout_ns == 0: wait for ever.
__global__ void ctc_reduce_loss_allreduce_kernel(const float* __restrict__ nll,
                                                 const int32_t* __restrict__ tgt_lens, int N, int mode,
                                                 const float* pair_in,
                                                 PeerBufs peers, int rank, int world, unsigned seq,
                                                 unsigned long long timeout_ns,
                                                 float* __restrict__ out2, float* __restrict__ loss,
                                                 int* __restrict__ status) {
    __shared__ double s_part[32];
    __shared__ float s_pair[2];
    __shared__ float2 s_in[8];
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (pair_in != nullptr) {
        if (threadIdx.x < 2) s_pair[threadIdx.x] = pair_in[threadIdx.x];
    } else {
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
            s_pair[0] = (float)s;
            s_pair[1] = (float)N;
        }
    }
    __syncthreads();
    const int par = (int)(seq & 1u);
    if ((int)threadIdx.x < world) {
        // one 16-byte store per peer, each 8-byte half carrying its own copy of the sequence number
        // (8 bytes is the unit NVLink delivers atomically: the NCCL "LL" convention)
        float4 v = make_float4(s_pair[0], __uint_as_float(seq), s_pair[1], __uint_as_float(seq));
        float4* dst = peers.p[threadIdx.x] + par * 8 + rank;
        asm volatile("st.volatile.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        __threadfence_system();
        // the pair rank `threadIdx.x` addressed to me
        const float4* src = peers.p[rank] + par * 8 + threadIdx.x;
        float4 r;
        const unsigned long long t0 = global_ns();
        for (unsigned spins = 0;; ++spins) {
            asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(src) : "memory");
            if (__float_as_uint(r.y) == seq && __float_as_uint(r.w) == seq) break;
            // a missing peer must fail loudly (status bit + NaN loss), not hang the stream for ever
            if (timeout_ns != 0 && (spins & 1023u) == 1023u && global_ns() - t0 > timeout_ns) {
                atomicOr(status, kStatusPeerTimeout);
                r.x = CUDART_NAN_F;
                r.z = 0.f;
                break;
            }
            __nanosleep(spins < 64 ? 100 : 1000);
        }
        s_in[threadIdx.x] = make_float2(r.x, r.z);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0, n = 0.0;
        for (int r = 0; r < world; ++r) { s += (double)s_in[r].x; n += (double)s_in[r].y; }
        out2[0] = (float)s;
        out2[1] = (float)n;
        if (loss) loss[0] = (mode == 1) ? (float)(s / fmax(n, 1.0)) : (float)s;
    }
}

}  // namespace ctcb200
