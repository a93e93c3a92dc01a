// ctc_launch_log.cu -- instantiations and launchers of the two LOG-domain kernels: the warp-specialised
// pipe kernel (ctc_pipe.cuh; per-utterance fallback of the linear kernel, CTC_B200_KERNEL=p) and the
// generic kernel (ctc_kernels.cuh; fallback for vocabularies whose rows are not 16-byte aligned, and
// targets too long for the other two).
#include "ctc_launch.h"

namespace ctcb200 {

namespace {

using PipeKernel = void (*)(const PipeParams);
using GenericKernel = void (*)(const FusedParams);

// Only the (P, CTA size) combinations the geometry choice of ctc_abi.cu can produce are instantiated:
// pipe:    P = 1 (<= 32 pairs) and P = 2 (<= 64 pairs) always fit 128 threads; P = 4 covers the rest
// generic: P = 1 up to 1024 pairs, P = 2 up to 2048, P = 4 up to 4096 (always more than 256 threads)
constexpr int kPipeCount = 7;
constexpr int kGenericCount = 4;

PipeKernel pipe_kernel(int id) {
    switch (id) {
        case 0: return ctc_pipe_kernel<1, 128, 4>;
        case 1: return ctc_pipe_kernel<2, 128, 4>;
        case 2: return ctc_pipe_kernel<4, 128, 4>;
        case 3: return ctc_pipe_kernel<4, 160, 4>;
        case 4: return ctc_pipe_kernel<4, 256, 2>;
        case 5: return ctc_pipe_kernel<4, 512, 1>;
        case 6: return ctc_pipe_kernel<4, 1024, 1>;
    }
    return nullptr;
}

GenericKernel generic_kernel(int id) {
    switch (id) {
        case 0: return ctc_fused_kernel<1, 256>;
        case 1: return ctc_fused_kernel<1, 1024>;
        case 2: return ctc_fused_kernel<2, 1024>;
        case 3: return ctc_fused_kernel<4, 1024>;
    }
    return nullptr;
}

const char* const kPipeNames[kPipeCount] = {
    "ctc_pipe_kernel<1,128,4>", "ctc_pipe_kernel<2,128,4>", "ctc_pipe_kernel<4,128,4>", "ctc_pipe_kernel<4,160,4>",
    "ctc_pipe_kernel<4,256,2>", "ctc_pipe_kernel<4,512,1>", "ctc_pipe_kernel<4,1024,1>",
};
const char* const kGenericNames[kGenericCount] = {
    "ctc_fused_kernel<1,256>", "ctc_fused_kernel<1,1024>", "ctc_fused_kernel<2,1024>", "ctc_fused_kernel<4,1024>",
};

SmemMark g_pipe_marks[kPipeCount], g_generic_marks[kGenericCount];

}  // namespace

int pipe_variant(const Geometry& g) {
    if (g.NT > 1024) return -1;
    if (g.P == 1) return g.NT <= 128 ? 0 : -1;
    if (g.P == 2) return g.NT <= 128 ? 1 : -1;
    if (g.P != 4) return -1;
    return g.NT <= 128 ? 2 : (g.NT <= 160 ? 3 : (g.NT <= 256 ? 4 : (g.NT <= 512 ? 5 : 6)));
}
const char* pipe_variant_name(int id) { return id >= 0 && id < kPipeCount ? kPipeNames[id] : "?"; }

cudaError_t launch_pipe(const PipeParams& pp, const Geometry& g, int n_utt, bool pdl, cudaStream_t st) {
    const int id = pipe_variant(g);
    PipeKernel k = pipe_kernel(id);
    if (!k) return last_cuda_error_set(cudaErrorInvalidConfiguration);
    cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k), g_pipe_marks[id], g.smem);
    if (e != cudaSuccess) return e;
    if (pdl)   // fallback pass right behind the linear kernel
        return last_cuda_error_set(launch_pdl(k, dim3(2 * n_utt), dim3(g.NT), (size_t)g.smem, st, pp));
    k<<<dim3(2 * n_utt), dim3(g.NT), g.smem, st>>>(pp);
    return last_cuda_error_set(cudaGetLastError());
}

int generic_variant(const Geometry& g) {
    if (g.NT > 1024) return -1;
    // small CTAs get the full register file, large ones the 64-register cap
    if (g.P == 1) return g.NT <= 256 ? 0 : 1;
    if (g.P == 2) return 2;
    return g.P == 4 ? 3 : -1;
}
const char* generic_variant_name(int id) { return id >= 0 && id < kGenericCount ? kGenericNames[id] : "?"; }

cudaError_t launch_generic(const FusedParams& prm, const Geometry& g, int n_utt, cudaStream_t st) {
    const int id = generic_variant(g);
    GenericKernel k = generic_kernel(id);
    if (!k) return last_cuda_error_set(cudaErrorInvalidConfiguration);
    cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k), g_generic_marks[id], g.smem);
    if (e != cudaSuccess) return e;
    k<<<dim3(2 * n_utt), dim3(g.NT), g.smem, st>>>(prm);
    return last_cuda_error_set(cudaGetLastError());
}

}  // namespace ctcb200
