// ctc_launch.h -- internal interface between the C ABI (ctc_abi.cu) and the translation units that
// instantiate the kernels (ctc_launch_lin.cu: ctc_lin.cuh; ctc_launch_log.cu: ctc_pipe.cuh and
// ctc_kernels.cuh's ctc_fused_kernel; ctc_decode.cu).  Split so that the instantiations compile in
// parallel; nothing here is exported.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>

#include "ctc_kernels.cuh"
#include "ctc_pipe.cuh"
#include "../../include/ctc_b200.h"

namespace ctcb200 {

struct Geometry {
    int pipe;       // 2: linear-domain kernel (ctc_lin.cuh) with a log-domain kernel as its per-utterance
                    //    fallback, 1: log-domain pipe kernel (ctc_pipe.cuh), 0: generic kernel (ctc_kernels.cuh)
    int base;       // which log-domain kernel the fields below describe: 1 pipe, 0 generic
    int P, NT, W, NP, chunk, RS, smem;
    int R, G, D;    // pipe only: recursion / gradient warps, fetch distance in chunks
    size_t lat_utt_stride;  // floats
    // pipe == 2: geometry of the linear kernel (the fields above describe the fallback)
    int lP, lNT, lNP, lchunk, lRS, lsmem, lR, lH, lD, lYS;
    int lriss;      // headline class, at most two CTAs per SM: the instantiation whose recursion warp requests the partner's rows
    size_t l_lat_utt_stride;
    size_t lattice_floats_per_utt() const {
        return pipe == 2 ? (lat_utt_stride > l_lat_utt_stride ? lat_utt_stride : l_lat_utt_stride) : lat_utt_stride;
    }
};

// Knobs read ONCE from the environment (developer tuning; the defaults are the product).
struct Env {
    int chunk, dist, rotate, utt_rot, nofix, pdl, slice_streams, persist, map_mode, helpers, rec_iss;
    char kernel;   // 'g': generic, 'p': log-domain pipe, 0: default (linear)
    static int geti(const char* name, int dflt) {
        const char* e = std::getenv(name);
        return e ? std::atoi(e) : dflt;
    }
    Env() {
        chunk = geti("CTC_B200_CHUNK", 0);
        dist = geti("CTC_B200_DIST", 0);
        rotate = geti("CTC_B200_ROTATE", 1);
        utt_rot = geti("CTC_B200_UTT_ROT", -1);
        map_mode = geti("CTC_B200_MAP", 0);
        rec_iss = geti("CTC_B200_REC_ISS", -1);   // developer knob: -1 automatic, 0 / 1 force
        helpers = geti("CTC_B200_HELPERS", 0);
        nofix = geti("CTC_B200_NOFIX", 0);
        pdl = geti("CTC_B200_PDL", 1);
        slice_streams = geti("CTC_B200_SLICE_STREAMS", 1);
        persist = geti("CTC_B200_PERSIST", -1);   // -1: automatic, 0: never, 1: always
        const char* k = std::getenv("CTC_B200_KERNEL");
        kernel = k ? k[0] : 0;
    }
};
const Env& env();

// Per-device facts (SM count) and per-(instantiation, device) "largest dynamic shared memory configured
// so far" marks: cudaFuncSetAttribute applies to the CURRENT device only, and one process may drive
// several devices (the torch shim switches with a CUDAGuard).
constexpr int kMaxDevices = 64;
int current_device();
int num_sms();
struct SmemMark {
    std::atomic<int> v[kMaxDevices];
    SmemMark() { for (auto& x : v) x.store(-1); }
};
// returns cudaSuccess when `bytes` of dynamic shared memory are configured for `func` on this device
cudaError_t ensure_smem(const void* func, SmemMark& mark, int bytes);

cudaError_t last_cuda_error_set(cudaError_t e);   // remembers e (thread-local) and returns it

// Linear-domain kernel.  `n_clusters` clusters are launched; pp.queue != nullptr makes them persistent.
// *variant receives the instantiation id (valid also when only asking: launch == false).
int lin_variant(const Geometry& g, int V);
int lin_smem_size(int NP, int R, int V, int TC, int RS, int ys);   // LinSmem(...).total
int lin_row_stride_host(int NP, int P);
const char* lin_variant_name(int id);
cudaError_t launch_lin(const PipeParams& pp, int* flags, const Geometry& g, int n_clusters, cudaStream_t st);
// co-resident clusters of the instantiation `g` selects, on the current device (0 when it cannot be queried)
int lin_resident_clusters(const Geometry& g, int V);
// persistent launches (utterance queue) exist for the headline shape class only
bool lin_supports_queue(const Geometry& g, int V);

int pipe_variant(const Geometry& g);
const char* pipe_variant_name(int id);
cudaError_t launch_pipe(const PipeParams& pp, const Geometry& g, int n_utt, bool pdl, cudaStream_t st);

int generic_variant(const Geometry& g);
const char* generic_variant_name(int id);
cudaError_t launch_generic(const FusedParams& prm, const Geometry& g, int n_utt, cudaStream_t st);

// Programmatic dependent launch for the small kernels that follow the fused kernel (its fallback pass
// and the loss reduction): the launch is set up while the previous kernel drains; the kernels
// themselves wait for it with griddepcontrol.wait before their first global-memory read.
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

// greedy decode + label error count (ctc_decode.cu)
cudaError_t launch_decode_ler(const float* acts, int T, int N, int V, long long frame_stride,
                              long long utt_stride, const int32_t* in_lens, const int32_t* targets,
                              const int32_t* tgt_off, const int32_t* tgt_lens, int blank, int32_t* hyp,
                              int32_t* hyp_len, int32_t* dist, long long* totals, cudaStream_t st);

}  // namespace ctcb200
