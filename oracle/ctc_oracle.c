/*
 * oracle/ctc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, double-precision restatement of the CTC loss path that
 * jinserk/pytorch-asr executes at asr/models/trainer.py:153 (construction of
 * nn.CTCLoss(blank=0, reduction='mean')) and asr/models/trainer.py:422 / :508
 * (the call `self.loss(ys_hat, ys, frame_lens, label_lens)`), preceded by the
 * network's nn.LogSoftmax at asr/models/deepspeech_ctc/network.py:375,395.
 *
 * The arithmetic itself is not in the reference tree: it is PyTorch's ATen
 * `ctc_loss` (un-vendored, un-pinned dependency; torch 2.11.0+cu128 in this
 * image).  This file restates the published algorithm (Graves et al. 2006,
 * the paper torch/nn/modules/loss.py cites for CTCLoss) with torch's
 * conventions (torch/nn/functional.py ctc_loss docstring):
 *   - log_probs are T x N x V, targets are a 1-D concatenation, blank index
 *     given, per-utterance input/target lengths;
 *   - nll_b = -log p(l_b | x_b); infeasible alignment => +inf;
 *   - gradient w.r.t. the *logits* (log_softmax folded in):
 *       grad[t,b,v] = scale_b * (softmax(x)[t,b,v] - occupancy[t,b,v])  t <  T_b
 *       grad[t,b,v] = 0                                                 t >= T_b
 *     where occupancy[t,b,v] = sum_{s: l'_s = v} alpha[t,s]*beta[t,s] /
 *     (y[t,l'_s] * p(l|x)), all evaluated in log space.
 *
 * Parity pin: the reference has no tests, fixtures or golden vectors (SURVEY.md
 * section 4, 8c), so this oracle is pinned against the reference's own
 * dependency run live -- torch.nn.functional.ctc_loss on CPU in fp64 and fp32
 * (tests/test_oracle.py) -- and against the committed fixtures in
 * tests/golden/ that were generated from it (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * call into this file.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/libctc_oracle.so oracle/ctc_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline double lse2(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    double m = a > b ? a : b;
    return m + log(exp(a - m) + exp(b - m));
}

static inline double lse3(double a, double b, double c) { return lse2(lse2(a, b), c); }

/* blank-extended label l'_s: even s -> blank, odd s -> targets[(s-1)/2] */
static inline int ext_label(const int32_t *tg, int s, int blank) {
    return (s & 1) ? tg[s >> 1] : blank;
}

/*
 * One utterance.  x: pointer to acts[0, b, 0]; consecutive frames are
 * `frame_stride` floats apart.  lp/alpha/beta are caller-provided scratch of
 * Tb*V, Tb*L, Tb*L doubles.  grad (may be NULL) is written for ALL T frames
 * of this utterance with the same frame stride, in doubles.
 * alpha_out (may be NULL): [T][Lmax] slab for this utterance, -inf padded,
 * matching the layout torch._ctc_loss returns (log_alpha[b, t, s]).
 */
static double one_utt(const float *x, size_t frame_stride, int T, int V, int Tb, int S,
                      const int32_t *tg, int blank, double scale, double *lp, double *alpha,
                      double *beta, double *grad, double *alpha_out, int Lmax) {
    const int L = 2 * S + 1;
    double nll;
    if (alpha_out)
        for (long k = 0; k < (long)T * Lmax; ++k) alpha_out[k] = -INFINITY;

    /* log_softmax over the vocabulary, per frame (network.py:375) */
    for (int t = 0; t < Tb; ++t) {
        const float *xr = x + (size_t)t * frame_stride;
        double m = -INFINITY, z = 0.0;
        for (int v = 0; v < V; ++v)
            if ((double)xr[v] > m) m = (double)xr[v];
        for (int v = 0; v < V; ++v) z += exp((double)xr[v] - m);
        double lz = m + log(z);
        for (int v = 0; v < V; ++v) lp[(size_t)t * V + v] = (double)xr[v] - lz;
    }

    if (Tb == 0) {
        /* torch: empty input => 0 for an empty target, +inf otherwise */
        nll = (S == 0) ? 0.0 : INFINITY;
    } else {
        /* alpha recursion */
        for (int s = 0; s < L; ++s) alpha[s] = -INFINITY;
        alpha[0] = lp[blank];
        if (S > 0) alpha[1] = lp[tg[0]];
        for (int t = 1; t < Tb; ++t) {
            const double *ap = alpha + (size_t)(t - 1) * L;
            double *ac = alpha + (size_t)t * L;
            const double *lpt = lp + (size_t)t * V;
            for (int s = 0; s < L; ++s) {
                int c = ext_label(tg, s, blank);
                double a = ap[s];
                double b = s >= 1 ? ap[s - 1] : -INFINITY;
                double d = (s >= 2 && (s & 1) && ext_label(tg, s - 2, blank) != c) ? ap[s - 2]
                                                                                  : -INFINITY;
                ac[s] = lse3(a, b, d) + lpt[c];
            }
        }
        const double *al = alpha + (size_t)(Tb - 1) * L;
        double ll = S > 0 ? lse2(al[L - 1], al[L - 2]) : al[L - 1];
        nll = -ll;
        if (alpha_out)
            for (int t = 0; t < Tb; ++t)
                memcpy(alpha_out + (size_t)t * Lmax, alpha + (size_t)t * L, sizeof(double) * L);
    }

    if (!grad) return nll;

    /* zero rows beyond the utterance length (torch writes exact zeros) */
    for (int t = Tb; t < T; ++t) {
        double *gr = grad + (size_t)t * frame_stride;
        for (int v = 0; v < V; ++v) gr[v] = 0.0;
    }
    if (Tb == 0) return nll;

    /* beta recursion */
    {
        double *bl = beta + (size_t)(Tb - 1) * L;
        const double *lpt = lp + (size_t)(Tb - 1) * V;
        for (int s = 0; s < L; ++s) bl[s] = -INFINITY;
        bl[L - 1] = lpt[blank];
        if (S > 0) bl[L - 2] = lpt[tg[S - 1]];
    }
    for (int t = Tb - 2; t >= 0; --t) {
        const double *bn = beta + (size_t)(t + 1) * L;
        double *bc = beta + (size_t)t * L;
        const double *lpt = lp + (size_t)t * V;
        for (int s = 0; s < L; ++s) {
            int c = ext_label(tg, s, blank);
            double a = bn[s];
            double b = s + 1 < L ? bn[s + 1] : -INFINITY;
            double d = (s + 2 < L && (s & 1) && ext_label(tg, s + 2, blank) != c) ? bn[s + 2]
                                                                                  : -INFINITY;
            bc[s] = lse3(a, b, d) + lpt[c];
        }
    }

    /* gradient: softmax - occupancy.  With nll = +inf this evaluates
     * exp(-inf + inf) = NaN, which is what torch returns as well. */
    for (int t = 0; t < Tb; ++t) {
        double *gr = grad + (size_t)t * frame_stride;
        const double *lpt = lp + (size_t)t * V;
        const double *ac = alpha + (size_t)t * L;
        const double *bc = beta + (size_t)t * L;
        for (int v = 0; v < V; ++v) gr[v] = -INFINITY; /* log-occupancy accumulators */
        for (int s = 0; s < L; ++s) {
            int c = ext_label(tg, s, blank);
            gr[c] = lse2(gr[c], ac[s] + bc[s]);
        }
        for (int v = 0; v < V; ++v) {
            double occ = exp(gr[v] + nll - lpt[v]);
            gr[v] = scale * (exp(lpt[v]) - occ);
        }
    }
    return nll;
}

/*
 * acts         [T, N, V] float32 raw logits (the engine fuses log_softmax)
 * targets      concatenated int32 labels; utterance b owns
 *              targets[tgt_offsets[b] .. tgt_offsets[b] + tgt_lens[b])
 * grad_scale   per-utterance multiplier (NULL => 1.0)
 * nll          [N] double out
 * grad         [T, N, V] double out, or NULL
 * alpha_out    [N, T, 2*Smax+1] double out (natural-log alpha, -inf padded), or NULL
 * returns 0, or -1 on allocation failure, or -2 on invalid lengths/labels.
 */
int ctc_oracle_f64(const float *acts, const int32_t *targets, const int32_t *tgt_offsets,
                   const int32_t *in_lens, const int32_t *tgt_lens, int T, int N, int V,
                   int blank, const double *grad_scale, double *nll, double *grad,
                   double *alpha_out, int Smax) {
    int status = 0;
    for (int b = 0; b < N; ++b) {
        if (in_lens[b] < 0 || in_lens[b] > T || tgt_lens[b] < 0) return -2;
        for (int k = 0; k < tgt_lens[b]; ++k) {
            int c = targets[tgt_offsets[b] + k];
            if (c < 0 || c >= V) return -2;
        }
    }
    const int Lmax = 2 * Smax + 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < N; ++b) {
        int Tb = in_lens[b], S = tgt_lens[b], L = 2 * S + 1;
        size_t nlat = (size_t)(Tb > 0 ? Tb : 1) * L;
        double *lp = (double *)malloc(sizeof(double) * (size_t)(Tb > 0 ? Tb : 1) * V);
        double *alpha = (double *)malloc(sizeof(double) * nlat);
        double *beta = grad ? (double *)malloc(sizeof(double) * nlat) : NULL;
        if (!lp || !alpha || (grad && !beta)) {
            status = -1;
        } else {
            nll[b] = one_utt(acts + (size_t)b * V, (size_t)N * V, T, V, Tb, S,
                             targets + tgt_offsets[b], blank, grad_scale ? grad_scale[b] : 1.0,
                             lp, alpha, beta, grad ? grad + (size_t)b * V : NULL,
                             alpha_out ? alpha_out + (size_t)b * T * Lmax : NULL, Lmax);
        }
        free(lp);
        free(alpha);
        free(beta);
    }
    return status;
}
