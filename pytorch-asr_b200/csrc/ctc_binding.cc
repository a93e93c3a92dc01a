// ctc_binding.cc -- torch C++ extension shim over the C ABI of libctc_b200.so.
//
// Built as module `torch_asr._ctc_lib` by pytorch-asr_b200/setup.py the way the
// reference builds `torch_asr._latgen_lib` (asr/kaldi/setup.py:48-71,
// PYBIND11_MODULE at asr/kaldi/src/latgen_lib.cc:278-281).  This file holds no
// CUDA code: tensors in, raw pointers + sizes + current stream out.
//
// Input conventions are the reference's call site (asr/models/trainer.py:409-444,
// asr/utils/dataloader.py:51-74): acts CUDA fp32 [T,N,V] contiguous; targets
// int32 1-D concatenated on the CPU; frame/label lengths int32 [N] on the CPU.
// Like torch.nn.functional.ctc_loss it also takes int64, CUDA-resident and 2-D
// padded targets, and CUDA-resident lengths (those cost a device->host sync, as
// they do in torch).
#include <torch/extension.h>

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ctc_b200.h"

namespace {

void check_status(int rc, const char* what) {
    if (rc == CTC_B200_OK) return;
    std::string msg = std::string(what) + ": " + ctc_b200_status_string(rc);
    if (rc == CTC_B200_CUDA_ERROR) msg += std::string(" (") + ctc_b200_last_cuda_error() + ")";
    TORCH_CHECK(false, msg.c_str());
}

// Error messages are formatted with vsnprintf and handed to TORCH_CHECK as ONE const char*:
// in this toolchain any std::ostringstream instantiated inside an extension translation
// unit that includes the torch headers crashes (c10::str goes through one), so the
// multi-argument form of TORCH_CHECK must not be used here.
[[noreturn]] void fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    TORCH_CHECK(false, buf);
    throw std::runtime_error(buf);  // not reached
}

size_t up16(size_t x) { return (x + 15) / 16 * 16; }

std::vector<int64_t> lengths_to_host(const torch::Tensor& t, int64_t N, const char* name) {
    if (!(t.dim() == 1 && t.size(0) == N)) fail("%s must have shape [N=%lld]", name, (long long)N);
    if (!(t.scalar_type() == torch::kInt32 || t.scalar_type() == torch::kInt64))
        fail("%s must be int32 or int64", name);
    torch::Tensor c = t.to(torch::kCPU, torch::kInt64).contiguous();  // syncs iff CUDA resident
    const int64_t* p = c.data_ptr<int64_t>();
    return std::vector<int64_t>(p, p + N);
}

// forward(acts, targets, input_lengths, target_lengths, blank, reduction, zero_infinity, want_grad,
//         batch_major, use_clamp, clamp_min, clamp_max, zero_on_short)
//   -> (loss, nll[N], grad (acts' shape) or empty, out2[2], status4[4])
// reduction: 0 none, 1 mean, 2 sum.  grad already carries the reduction's scale
// (1/(N*max(S_b,1)) for mean), i.e. it is d loss / d acts for grad_output == 1.
// batch_major: acts (and grad) are [N,T,V], the network's own layout (folds trainer.py:418).
// use_clamp: fused Hardtanh(clamp_min, clamp_max) (network.py:370); grad is then w.r.t. the raw logits.
// status4 = [loss, flags (int bits: 1 NaN, 2 inf, 4 some T_b < 2 S_b), backward factor, #short]
// (trainer.py:423-430 in one read); zero_on_short applies the reference's loss.mul_(0) on the device.
std::vector<torch::Tensor> forward(const torch::Tensor& acts, const torch::Tensor& targets,
                                   const torch::Tensor& input_lengths,
                                   const torch::Tensor& target_lengths, int64_t blank,
                                   int64_t reduction, bool zero_infinity, bool want_grad,
                                   bool batch_major, bool use_clamp, double clamp_min, double clamp_max,
                                   bool zero_on_short) {
    TORCH_CHECK(acts.is_cuda(), "ctc_b200: acts must be a CUDA tensor (no CPU fallback)");
    TORCH_CHECK(acts.scalar_type() == torch::kFloat32, "ctc_b200: acts must be float32");
    TORCH_CHECK(acts.dim() == 3, "ctc_b200: acts must be [T, N, V] (or [N, T, V] with batch_major)");
    TORCH_CHECK(reduction >= 0 && reduction <= 2, "ctc_b200: bad reduction");
    TORCH_CHECK(!use_clamp || clamp_min < clamp_max, "ctc_b200: clamp_min must be < clamp_max");
    torch::Tensor x = acts.contiguous();
    const int64_t T = batch_major ? x.size(1) : x.size(0), N = batch_major ? x.size(0) : x.size(1), V = x.size(2);
    TORCH_CHECK(blank >= 0 && blank < V, "ctc_b200: blank must be in label range");
    TORCH_CHECK(T < (1LL << 30) && N < (1LL << 30) && V < (1LL << 30), "ctc_b200: size overflow");

    const std::vector<int64_t> il = lengths_to_host(input_lengths, N, "input_lengths");
    const std::vector<int64_t> tl = lengths_to_host(target_lengths, N, "target_lengths");
    int64_t S_max = 0, total = 0;
    for (int64_t b = 0; b < N; ++b) {
        if (!(il[b] >= 0 && il[b] <= T))
            fail("Expected input_lengths to have value at most %lld, but got value %lld", (long long)T,
                 (long long)il[b]);
        if (tl[b] < 0)
            fail("Expected target_lengths to have value at least 0, but got value %lld", (long long)tl[b]);
        S_max = std::max(S_max, tl[b]);
        total += tl[b];
    }
    TORCH_CHECK(targets.scalar_type() == torch::kInt32 || targets.scalar_type() == torch::kInt64,
                "ctc_b200: targets must be int32 or int64");
    const bool padded = targets.dim() == 2;
    TORCH_CHECK(padded || targets.dim() == 1, "ctc_b200: targets must be 1-D or 2-D");
    if (padded) {
        TORCH_CHECK(targets.size(0) == N && targets.size(1) >= S_max,
                    "ctc_b200: padded targets must be [N, >= max target length]");
    } else {
        TORCH_CHECK(targets.size(0) >= total, "ctc_b200: concatenated targets shorter than sum(target_lengths)");
    }

    const bool tg_on_host = !targets.is_cuda();
    if (tg_on_host) {   // host-resident labels (the reference's case) are validated here, for free
        torch::Tensor tc = targets.contiguous();
        auto check = [&](auto* src) {
            const int64_t stride = padded ? tc.size(1) : 0;
            for (int64_t b = 0, k = 0; b < N; ++b)
                for (int64_t j = 0; j < tl[b]; ++j, ++k) {
                    const int64_t c = padded ? (int64_t)src[b * stride + j] : (int64_t)src[k];
                    if (!(c >= 0 && c < V))
                        fail("ctc_b200: target label %lld outside [0, %lld)", (long long)c, (long long)V);
                }
        };
        if (tc.scalar_type() == torch::kInt32) check(tc.data_ptr<int32_t>());
        else check(tc.data_ptr<int64_t>());
    }
    // every argument check is done; from here on work is enqueued on acts' device
    const c10::cuda::CUDAGuard guard(acts.device());

    // ---- one pinned staging block: [targets | offsets | in_lens | tgt_lens | scale] ----
    const size_t o_tg = 0;
    const size_t o_off = up16((size_t)std::max<int64_t>(total, 1) * 4);
    const size_t o_il = o_off + up16((size_t)N * 4);
    const size_t o_tl = o_il + up16((size_t)N * 4);
    const size_t o_sc = o_tl + up16((size_t)N * 4);
    const size_t bytes = o_sc + up16((size_t)N * 4);
    torch::Tensor h = torch::empty({(int64_t)bytes},
                                   torch::TensorOptions().dtype(torch::kUInt8).pinned_memory(true));
    char* hp = static_cast<char*>(h.data_ptr());
    int32_t* h_tg = reinterpret_cast<int32_t*>(hp + o_tg);
    int32_t* h_off = reinterpret_cast<int32_t*>(hp + o_off);
    int32_t* h_il = reinterpret_cast<int32_t*>(hp + o_il);
    int32_t* h_tl = reinterpret_cast<int32_t*>(hp + o_tl);
    float* h_sc = reinterpret_cast<float*>(hp + o_sc);
    int64_t acc = 0;
    for (int64_t b = 0; b < N; ++b) {
        h_off[b] = (int32_t)acc;
        acc += tl[b];
        h_il[b] = (int32_t)il[b];
        h_tl[b] = (int32_t)tl[b];
        h_sc[b] = reduction == 1 ? 1.0f / ((float)N * (float)std::max<int64_t>(tl[b], 1)) : 1.0f;
    }
    torch::Tensor tg_dev;  // only used when the targets are CUDA resident
    if (tg_on_host) {
        torch::Tensor tc = targets.contiguous();
        auto put = [&](auto* src) {
            if (padded) {
                const int64_t stride = tc.size(1);
                int64_t k = 0;
                for (int64_t b = 0; b < N; ++b)
                    for (int64_t j = 0; j < tl[b]; ++j) h_tg[k++] = (int32_t)src[b * stride + j];
            } else {
                for (int64_t k = 0; k < total; ++k) h_tg[k] = (int32_t)src[k];
            }
        };
        if (tc.scalar_type() == torch::kInt32) put(tc.data_ptr<int32_t>());
        else put(tc.data_ptr<int64_t>());
    } else {
        // CUDA-resident targets: pack on the device (rare path; torch syncs here as well)
        if (padded) {
            std::vector<torch::Tensor> parts;
            for (int64_t b = 0; b < N; ++b) parts.push_back(targets[b].narrow(0, 0, tl[b]));
            tg_dev = torch::cat(parts).to(torch::kInt32).contiguous();
        } else {
            tg_dev = targets.narrow(0, 0, total).to(torch::kInt32).contiguous();
        }
        if (tg_dev.numel() == 0) tg_dev = torch::zeros({1}, tg_dev.options());
    }
    torch::Tensor d = h.to(x.device(), /*non_blocking=*/true);
    char* dp = static_cast<char*>(d.data_ptr());
    const int32_t* d_tg = tg_on_host ? reinterpret_cast<const int32_t*>(dp + o_tg)
                                     : tg_dev.data_ptr<int32_t>();
    const int32_t* d_off = reinterpret_cast<const int32_t*>(dp + o_off);
    const int32_t* d_il = reinterpret_cast<const int32_t*>(dp + o_il);
    const int32_t* d_tl = reinterpret_cast<const int32_t*>(dp + o_tl);
    const float* d_sc = reinterpret_cast<const float*>(dp + o_sc);

    auto fopt = x.options();
    torch::Tensor nll = torch::empty({N}, fopt);
    torch::Tensor grad = want_grad ? torch::empty_like(x) : torch::empty({0}, fopt);
    torch::Tensor out2 = torch::empty({2}, fopt);
    torch::Tensor status4 = torch::empty({4}, fopt);
    torch::Tensor loss = torch::empty({}, fopt);
    ctc_b200_options opt;
    opt.layout = batch_major ? CTC_B200_LAYOUT_NTV : CTC_B200_LAYOUT_TNV;
    opt.use_clamp = use_clamp ? 1 : 0;
    opt.clamp_min = (float)clamp_min;
    opt.clamp_max = (float)clamp_max;
    opt.persistent = 0;

    cudaStream_t stream = at::cuda::getCurrentCUDAStream();
    if (N > 0) {
        size_t ws_bytes = 0;
        check_status(ctc_b200_workspace_bytes((int)T, (int)N, (int)V, (int)S_max, &ws_bytes),
                     "ctc_b200_workspace_bytes");
        torch::Tensor ws = torch::empty({(int64_t)ws_bytes},
                                        torch::TensorOptions().dtype(torch::kUInt8).device(x.device()));
        check_status(ctc_b200_clear_status(ws.data_ptr(), stream), "ctc_b200_clear_status");
        check_status(ctc_b200_fwd_bwd_ex_f32(x.data_ptr<float>(), d_tg, d_off, d_il, d_tl, (int)T,
                                             (int)N, (int)V, (int)S_max, (int)blank,
                                             zero_infinity ? 1 : 0, 0, (int)N, nll.data_ptr<float>(),
                                             want_grad ? grad.data_ptr<float>() : nullptr, d_sc,
                                             ws.data_ptr(), ws_bytes, &opt, stream),
                     "ctc_b200_fwd_bwd_ex_f32");
        // CUDA-resident targets could not be validated on the host: read the device-side check back
        // (one sync on this rare path; torch synchronises for CUDA-resident targets as well)
        if (!tg_on_host) check_status(ctc_b200_check_status(ws.data_ptr(), stream), "ctc_b200: targets");
    }
    // loss reduction + the trainer's post-loss checks (trainer.py:423-430) in ONE launch
    check_status(ctc_b200_reduce_loss_status_f32(nll.data_ptr<float>(), d_il, d_tl, (int)N,
                                                 reduction == 1 ? CTC_B200_REDUCE_MEAN : CTC_B200_REDUCE_SUM,
                                                 zero_on_short ? 1 : 0, out2.data_ptr<float>(),
                                                 loss.data_ptr<float>(), status4.data_ptr<float>(), stream),
                 "ctc_b200_reduce_loss_status_f32");
    // `d`, `ws` go back to the caching allocator here; it is stream-ordered, so the
    // kernels enqueued above still own them.
    return {loss, nll, grad, out2, status4};
}

// grad *= scale (scalar tensor, or [N] per utterance), in place; ==1 is free.
void scale_grad(torch::Tensor grad, const torch::Tensor& scale, bool batch_major) {
    TORCH_CHECK(grad.is_cuda() && grad.is_contiguous() && grad.dim() == 3 &&
                grad.scalar_type() == torch::kFloat32, "ctc_b200: bad grad tensor");
    const c10::cuda::CUDAGuard guard(grad.device());
    torch::Tensor s = scale.to(grad.device(), torch::kFloat32).contiguous();
    const int per_utt = s.numel() == 1 ? 0 : 1;
    const int64_t T = batch_major ? grad.size(1) : grad.size(0), N = batch_major ? grad.size(0) : grad.size(1);
    TORCH_CHECK(per_utt == 0 || s.numel() == N, "ctc_b200: scale must be scalar or [N]");
    check_status(ctc_b200_scale_grad_ex_f32(grad.data_ptr<float>(), s.data_ptr<float>(), per_utt,
                                            (int)T, (int)N, (int)grad.size(2),
                                            batch_major ? CTC_B200_LAYOUT_NTV : CTC_B200_LAYOUT_TNV,
                                            at::cuda::getCurrentCUDAStream()),
                 "ctc_b200_scale_grad_ex_f32");
}

std::vector<int64_t> geometry(int64_t T, int64_t N, int64_t V, int64_t S_max) {
    ctc_b200_geometry g;
    check_status(ctc_b200_get_geometry((int)T, (int)N, (int)V, (int)S_max, &g), "ctc_b200_get_geometry");
    return {g.kernel, g.rec_warps, g.grad_warps, g.pairs_per_thread, g.threads, g.chunk, g.row_stride, g.smem_bytes,
            (int64_t)g.workspace_bytes, g.variant, g.fallback_kernel, g.comb_groups, g.resident_clusters};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("forward", &forward, "fused log_softmax + CTC loss + gradient (B200)");
    m.def("scale_grad", &scale_grad, "apply autograd grad_output to the stored gradient");
    m.def("geometry", &geometry, "launch geometry for a problem size");
    m.def("version", []() { return ctc_b200_version(); });
}
