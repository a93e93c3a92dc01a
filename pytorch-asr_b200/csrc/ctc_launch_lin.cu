// ctc_launch_lin.cu -- instantiations of the linear-domain kernel (ctc_lin.cuh) and their launcher.
//
// One table entry per template instantiation; `lin_variant` is the single place that maps a geometry to
// an entry, so what ctc_b200_get_geometry reports is what is launched (tests/test_gpu_variants.py walks
// the table).
#include "ctc_launch.h"
#include "ctc_lin.cuh"

namespace ctcb200 {

namespace {

enum LinVariant {
    kLinFix = 0,        // <8,1,80,128,4,FIX>  V = 48, S <= 248: the headline shape class (C1, C2, C5)
    kLinR1Y80,          // <8,1,80,128,4>      48 < V <= 60 (V % 4 == 0), S <= 248, more than 74 utterances
    kLinR1,             // <8,1,0,128,4>       60 < V <= 256 or V % 4 != 0 (the reference's V = 177), S <= 248
    kLinR1Wide,         // <8,1,0,256,2>       V > 256 (C4), S <= 248
    kLinR2Y80,          // <8,2,80,512,1>      V <= 60, 249 <= S <= 504
    kLinR4Y80,          // <8,4,80,512,1>      V <= 60, 761 <= S <= 1016 (C3)
    kLinRn256,          // <8,0,0,256,2>       several recursion warps, run-time strides, <= 256 threads
    kLinRn512,          // <8,0,0,512,1>
    kLinRn1024,         // <8,0,0,1024,1>
    kLinR1Mid,          // <8,1,0,256,2,MID>   128 < V <= 256, or 60 < V <= 256 with rows that are not 16-byte aligned
                        // (the reference's V = 177): four helper warps
    kLinR1WideAl,       // <8,1,0,256,2,WIDE>  V > 256 in 16-byte aligned rows (C4); kLinR1Wide keeps the rows that are not
    kLinFixQueue,       // <8,1,80,128,4,FIX,QUEUE>  the headline shape class as a persistent launch (not reported as a
                        // variant of its own: same code, wrapped in the loop over the utterance queue)
    kLinR1Mid8,         // <8,1,0,512,1,MID>   the MID vocabularies in launches of at most one CTA per SM (the reference's
                        // batches of 32 / 64 at V = 177): eight helper warps, a warp per frame, and four copy warps
    kLinFixV,           // <8,1,80,128,4,FIX,VRUN>  the headline code for the narrower aligned vocabularies (V = 4 ... 44, V % 4 = 0)
    kLinFixRiss,        // <8,1,80,128,4,FIX,RISS>  the headline class in launches of at most two CTAs per SM (C1): the recursion
                        // warp requests the partner's rows
    kLinCount
};

const char* const kLinNames[kLinCount] = {
    "ctc_lin_kernel<8,1,80,128,4,FIX>", "ctc_lin_kernel<8,1,80,128,4>", "ctc_lin_kernel<8,1,0,128,4>",
    "ctc_lin_kernel<8,1,0,256,2>",      "ctc_lin_kernel<8,2,80,512,1>", "ctc_lin_kernel<8,4,80,512,1>",
    "ctc_lin_kernel<8,0,0,256,2>",      "ctc_lin_kernel<8,0,0,512,1>",  "ctc_lin_kernel<8,0,0,1024,1>",
    "ctc_lin_kernel<8,1,0,256,2,MID>", "ctc_lin_kernel<8,1,0,256,2,WIDE>", "ctc_lin_kernel<8,1,80,128,4,FIX,QUEUE>",
    "ctc_lin_kernel<8,1,0,512,1,MID>", "ctc_lin_kernel<8,1,80,128,4,FIX,VRUN>",
    "ctc_lin_kernel<8,1,80,128,4,FIX,RISS>",
};

using LinKernel = void (*)(const PipeParams, int*);

LinKernel lin_kernel(int id) {
    switch (id) {
        case kLinFix: return ctc_lin_kernel<8, 1, 80, 128, 4, true>;
        case kLinR1Y80: return ctc_lin_kernel<8, 1, 80, 128, 4>;
        case kLinR1: return ctc_lin_kernel<8, 1, 0, 128, 4>;
        case kLinR1Wide: return ctc_lin_kernel<8, 1, 0, 256, 2>;
        case kLinR2Y80: return ctc_lin_kernel<8, 2, 80, 512, 1>;
        case kLinR4Y80: return ctc_lin_kernel<8, 4, 80, 512, 1>;
        case kLinRn256: return ctc_lin_kernel<8, 0, 0, 256, 2>;
        case kLinRn512: return ctc_lin_kernel<8, 0, 0, 512, 1>;
        case kLinRn1024: return ctc_lin_kernel<8, 0, 0, 1024, 1>;
        case kLinR1Mid: return ctc_lin_kernel<8, 1, 0, 256, 2, false, false, true>;
        case kLinR1WideAl: return ctc_lin_kernel<8, 1, 0, 256, 2, false, false, false, true>;
        case kLinFixQueue: return ctc_lin_kernel<8, 1, 80, 128, 4, true, true>;
        case kLinR1Mid8: return ctc_lin_kernel<8, 1, 0, 512, 1, false, false, true>;
        case kLinFixV: return ctc_lin_kernel<8, 1, 80, 128, 4, true, false, false, false, true>;
        case kLinFixRiss: return ctc_lin_kernel<8, 1, 80, 128, 4, true, false, false, false, false, true>;
    }
    return nullptr;
}

SmemMark g_marks[kLinCount];

}  // namespace

int lin_variant(const Geometry& g, int V) {
    if (g.lP != 8) return -1;
    if (g.lR == 1) {
        if (g.lYS == 80 && g.lNT == 128) {
            if (g.lH == 1 && g.lD == 2 && !env().nofix && V % 4 == 0 && V >= 4 && V <= 48)
                return V == 48 ? (g.lriss ? kLinFixRiss : kLinFix) : kLinFixV;
            return kLinR1Y80;
        }
        if (g.lYS != 0) return -1;
        if (g.lNT <= 128) return kLinR1;
        // eight helpers + four copy warps, 480 threads: the MID vocabularies when every CTA has an SM of its own (pick_lin)
        if (g.lNT == 480) return (g.lH == 8 && g.lD == 2 && V <= 256 && g.lchunk == 4) ? kLinR1Mid8 : -1;
        if (g.lNT > 256) return -1;
        // (the WIDE / MID instantiations have their CTA shape -- four helpers, 224 threads, WIDE: chunks of 2 frames --
        // as compile-time constants; any other choice of the geometry heuristics runs the general instantiation)
        if (V > 256) return (V % 4 == 0 && g.lH == 4 && g.lD == 2 && g.lchunk == 2) ? kLinR1WideAl : kLinR1Wide;
        // MID keeps a frame's classes in registers, at most 16 per lane: at least 16 lanes per frame
        return (g.lH == 4 && g.lD == 2 && V <= 256 && g.lchunk == 4) ? kLinR1Mid : kLinR1Wide;
    }
    if (g.lYS == 80 && g.lNT <= 512 && g.lH == 2 && g.lD == 2) {   // (their CTA shape is a compile-time constant)
        if (g.lR == 2) return kLinR2Y80;
        if (g.lR == 4) return kLinR4Y80;
    }
    if (g.lYS != 0) return -1;   // (pick_lin only asks for the fixed emission-ring stride where an instantiation exists)
    if (g.lNT <= 256) return kLinRn256;
    if (g.lNT <= 512) return kLinRn512;
    return g.lNT <= 1024 ? kLinRn1024 : -1;
}

int lin_smem_size(int NP, int R, int V, int TC, int RS, int ys) { return LinSmem(NP, R, V, TC, RS, ys).total; }
int lin_row_stride_host(int NP, int P) { return lin_row_stride(NP, P); }

const char* lin_variant_name(int id) { return id >= 0 && id < kLinCount ? kLinNames[id] : "?"; }

bool lin_supports_queue(const Geometry& g, int V) { return lin_variant(g, V) == kLinFix; }

cudaError_t launch_lin(const PipeParams& pp, int* flags, const Geometry& g, int n_clusters, cudaStream_t st) {
    int id = lin_variant(g, pp.f.V);
    if (pp.queue != nullptr) {
        if (id != kLinFix) return last_cuda_error_set(cudaErrorInvalidConfiguration);   // (the caller asks lin_supports_queue)
        id = kLinFixQueue;
    }
    LinKernel k = lin_kernel(id);
    if (!k) return last_cuda_error_set(cudaErrorInvalidConfiguration);
    cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k), g_marks[id], g.lsmem);
    if (e != cudaSuccess) return e;
    k<<<dim3(2 * n_clusters), dim3(g.lNT), g.lsmem, st>>>(pp, flags);
    return last_cuda_error_set(cudaGetLastError());
}

int lin_resident_clusters(const Geometry& g, int V) {
    int id = lin_variant(g, V);
    if (id == kLinFix) id = kLinFixQueue;   // what a persistent launch would run
    LinKernel k = lin_kernel(id);
    if (!k) return 0;
    if (ensure_smem(reinterpret_cast<const void*>(k), g_marks[id], g.lsmem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    // co-resident clusters as the driver computes them for this kernel's (compile-time) 2-CTA clusters
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * num_sms() * 8);
    cfg.blockDim = dim3(g.lNT);
    cfg.dynamicSmemBytes = (size_t)g.lsmem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) == cudaSuccess && n > 0) return n;
    cudaGetLastError();
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, g.lNT, (size_t)g.lsmem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return per_sm * (num_sms() / 2);   // a 2-CTA cluster sits on one SM pair: whole pairs only
}

}  // namespace ctcb200
