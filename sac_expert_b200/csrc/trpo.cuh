// TRPO surrogate objective, its output cotangent, and the line-search statistics, on the rows bound as fvp_states
// (sac_eo/algs/model_free/trpo.py:36-63 surrogate + entropy regulariser, :229-317 back-tracking line search), in the
// GaussianActor._forward parameterisation (sac_eo/actors/continuous_actors.py:74-100; neglogp :137-143, entropy
// :145-148, kl :165-184, get_kl_info :186-192).  The MLP forward / VJP GEMMs are issued by saceo.cu on the Fisher-vector
// workspace (fvp.cuh); here are the per-row head, the fixed-order per-agent means and the parameter step.
#pragma once
#include "fvp.cuh"

namespace saceo {

// (mean, logstd, d logstd / d raw) of action dimension j at one state row
struct GaussRow { float mean, ls, dls; };
__device__ __forceinline__ GaussRow gauss_row(const KCtx& c, const float* out, const float* theta, int j, float ls_init,
                                              float floor_ls) {
  GaussRow g; g.mean = out[j];
  if (c.per_state_std) {
    const float o2 = out[c.A + j];
    const float sp = softplusf(o2);
    const float ls = logf(sp) + ls_init;
    g.dls = ls >= floor_ls ? (1.f / (1.f + expf(-o2))) / sp : 0.f;
    g.ls = fmaxf(ls, floor_ls);
  } else {
    const float ls = theta[c.L.na - c.A + j] + ls_init;
    g.dls = ls >= floor_ls ? 1.f : 0.f;
    g.ls = fmaxf(ls, floor_ls);
  }
  return g;
}

// per (agent, state row).  grid: (ceil(N/128), n_agents), block 128
//   nlp  = 0.5 sum_j(((a-mean)/exp(ls))^2 + 2 ls + log 2pi)         ent = 0.5 sum_j(2 ls + log 2pi + 1)
//   kl   = 0.5 sum_j(((mean-mean_ref)^2 + exp(2 ls_ref)) / exp(2 ls) + 2 ls - 2 ls_ref - 1)      (forward KL)
//   ratio = exp(nlp_old - nlp)      row statistics (f.Tmp[row*4..]) = { ratio adv, kl, |ratio - 1|, ent }
//   want_grad: G = d/d(raw outputs) of  mean(-ratio adv) - alpha (mean ent - ent_targ)   (clip_eps >= 0: PPO's clipped
//   surrogate instead of -ratio adv):
//     d/dmean_j = -adv ratio q_j / std_j / N,  d/dls_j = (adv ratio (1 - q_j^2) - alpha) / N,  q = (a-mean)/std
__global__ void k_trpo_rows(KCtx c, FvpWs f, const float* __restrict__ act, const float* __restrict__ adv,
                            const float* __restrict__ nlp_old, const float* __restrict__ kl_ref,
                            const float* __restrict__ alpha, float std_mult, int want_grad, float clip_eps,
                            float* __restrict__ nlp_out, float* __restrict__ kl_info_out) {
  const int agent = blockIdx.y;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= f.N) return;
  const int A = c.A;
  const long long r = (long long)agent * f.N + row;
  const float* out = f.Out + r * c.Ao;
  const float* theta = c.T.actor + (long long)agent * c.L.na_stride;
  const float ls_init = c.per_state_std ? logf(std_mult) - logf(kLog2) : logf(std_mult);
  const float floor_ls = logf(1e-3f);
  float nlp = 0.f, ent = 0.f, kl = 0.f;
  for (int j = 0; j < A; ++j) {
    const GaussRow g = gauss_row(c, out, theta, j, ls_init, floor_ls);
    const float q = act ? (act[r * A + j] - g.mean) / expf(g.ls) : 0.f;
    nlp += q * q + 2.f * g.ls + kLog2Pi;
    ent += 2.f * g.ls + kLog2Pi + 1.f;
    if (kl_ref) {
      const float mr = kl_ref[(r * A + j) * 2], lr = kl_ref[(r * A + j) * 2 + 1];
      const float dm = g.mean - mr;
      kl += (dm * dm + expf(2.f * lr)) / expf(2.f * g.ls) + 2.f * g.ls - 2.f * lr - 1.f;
    }
    if (kl_info_out) { kl_info_out[(r * A + j) * 2] = g.mean; kl_info_out[(r * A + j) * 2 + 1] = g.ls; }
  }
  nlp *= 0.5f; ent *= 0.5f; kl *= 0.5f;
  const float ratio = nlp_old ? expf(nlp_old[r] - nlp) : 1.f;
  const float av = adv ? adv[r] : 0.f;
  if (nlp_out) nlp_out[r] = nlp;
  float* rs = f.Tmp + r * 4;
  rs[0] = ratio * av; rs[1] = kl; rs[2] = fabsf(ratio - 1.f); rs[3] = ent;
  if (!want_grad) return;
  const float invN = 1.f / (float)f.N;
  float w = av * ratio * invN;
  if (clip_eps >= 0.f) {
    // PPO (ppo.py:135-141): mean(max(-ratio adv, -clip(ratio, 1-eps, 1+eps) adv)); tf.maximum sends the gradient to its
    // first argument when it is >= the second, and the clipped branch has zero gradient outside [1-eps, 1+eps]
    const float rc = fminf(fmaxf(ratio, 1.f - clip_eps), 1.f + clip_eps);
    if (!(-ratio * av >= -rc * av)) w = 0.f;
  }
  const float al = alpha ? alpha[agent] * invN : 0.f;
  float* g_out = f.G + r * c.Ao;
  for (int j = 0; j < A; ++j) {
    const GaussRow g = gauss_row(c, out, theta, j, ls_init, floor_ls);
    const float sd = expf(g.ls);
    const float q = (act[r * A + j] - g.mean) / sd;
    g_out[j] = -w * q / sd;
    const float gl = (w * (1.f - q * q) - al) * g.dls;
    if (c.per_state_std) g_out[A + j] = gl;
    else f.gls[r * A + j] = gl;
  }
}

// per-agent means of the four row statistics, fixed summation order.  grid: (n_agents), block 256
//   stats[agent*8 + {0: surr = mean(ratio adv), 1: kl, 2: tv = 0.5 mean|ratio-1|, 3: ent}]
__global__ void k_trpo_reduce(FvpWs f, float* __restrict__ stats) {
  __shared__ float sh[32];
  const int agent = blockIdx.x;
  const float* rs = f.Tmp + (long long)agent * f.N * 4;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int r = threadIdx.x; r < f.N; r += blockDim.x) {
    a0 += rs[r * 4]; a1 += rs[r * 4 + 1]; a2 += rs[r * 4 + 2]; a3 += rs[r * 4 + 3];
  }
  a0 = block_sum(a0, sh); a1 = block_sum(a1, sh); a2 = block_sum(a2, sh); a3 = block_sum(a3, sh);
  if (threadIdx.x == 0) {
    const float invN = 1.f / (float)f.N;
    float* s = stats + agent * 8;
    s[0] = a0 * invN; s[1] = a1 * invN; s[2] = 0.5f * a2 * invN; s[3] = a3 * invN;
    s[4] = 0.f; s[5] = 0.f; s[6] = 0.f; s[7] = 0.f;
  }
}

// actor.set_weights(theta_ref) ; actor.set_weights(scale * dir, from_flat=True, increment=True) with the floor of the
// state-independent logstd variable (continuous_actors.py:211-233).  grid: (ceil(na/256), n_agents)
__global__ void k_actor_step(KCtx c, const float* __restrict__ theta_ref, const float* __restrict__ dir,
                             const float* __restrict__ scale) {
  const int agent = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.L.na) return;
  const long long o = (long long)agent * c.L.na_stride + i;
  float v = __fadd_rn(theta_ref[o], __fmul_rn(scale[agent], dir[o]));
  if (!c.per_state_std && i >= c.L.na - c.A) v = fmaxf(v, logf(1e-3f));
  c.T.actor[o] = v;
}

// tf.clip_by_global_norm over the actor's trainable tensors (ppo.py:226-231): norm = sqrt(sum g^2) (fixed order),
// g *= max_norm / max(norm, max_norm); max_norm <= 0: norm only.  stats[agent*8 + {4: norm before, 5: norm after}].
// grid: (n_agents), block 256
__global__ void k_grad_clip(KCtx c, float* __restrict__ g, float max_norm, float* __restrict__ stats) {
  __shared__ float sh[32];
  const int agent = blockIdx.x;
  float* ga = g + (long long)agent * c.L.na_stride;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < c.L.na; i += blockDim.x) acc += ga[i] * ga[i];
  const float norm = sqrtf(block_sum(acc, sh));
  float post = norm;
  if (max_norm > 0.f) {
    const float scale = max_norm / fmaxf(norm, max_norm);
    for (long long i = threadIdx.x; i < c.L.na; i += blockDim.x) ga[i] *= scale;
    post = norm * scale;
  }
  if (threadIdx.x == 0 && stats) { stats[agent * 8 + 4] = norm; stats[agent * 8 + 5] = post; }
}

// ------------------------------------------------------------------------------------------
// Expert-observation term of the on-policy classes (trpo.py:92-158 two-model branch, ppo.py:176-213 with models[0]):
//   counterfactual a = GaussianActor.sample(sE) = mean + exp(logstd) u   (continuous_actors.py:103-123, _forward
//   parameterisation, NO squash; clip != 0: actor.tf_clip to the action limits, :128-129), through the frozen model(s),
//   MSE = mean_i 0.5 sum_j (s'E - pred)^2, gradient w.r.t. the actor's trainable variables.
// The E expert rows occupy rows [0, E) of the actor-phase buffers (Xpi, aOut, daOut, dls, Xm); the model term itself
// is the SAC-EO kernel (model_term.cuh) with the expert weight forced to 1.
// ------------------------------------------------------------------------------------------
// Xpi[i] = N_s(sE[perm i]);  Xm[net][il, :S] = N_s^M(sE[perm i]).  grid: (ceil(E*S/256), n_agents)
__global__ void k_exp_stage(KCtx c) {
  const int agent = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int S = c.S, SA = S + c.A;
  if (e >= c.E * S) return;
  const int i = e / S, j = e - i * S;
  const float* nr = c.T.norm + (long long)agent * c.L.norm_stride;
  const int src = c.perm[(long long)agent * c.E + i];
  const float x = c.expert_s[((long long)agent * c.E + src) * S + j];
  c.Xpi[((long long)agent * c.Rs + i) * c.ldXp + j] = (x - nr[c.L.off_s_mean + j]) / nstd(nr[c.L.off_s_std + j]);
  const int half = c.nmod == 2 ? c.E / 2 : c.E;
  const int net = i / half, il = i - net * half;
  c.Xm[(((long long)agent * 2 + net) * c.E + il) * SA + j] = (x - nr[c.L.off_m_s_mean + j]) / nstd(nr[c.L.off_m_s_std + j]);
}

// forward: action columns of Xm.  backward (bwd != 0): d(out) from mdXa.  One thread per expert row.
// noise: rows [2B, 2B + E) of the per-agent block (the u3 | u4 slots of the SAC-EO draw order).  grid: (ceil(E/128), n_agents)
__global__ void k_ghead(KCtx c, float std_mult, int clip, int bwd) {
  const int agent = blockIdx.y;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= c.E) return;
  const int A = c.A, Ao = c.Ao, S = c.S, SA = S + A, B = c.B, E = c.E;
  const float* out = c.aOut + ((long long)agent * c.Rs + row) * Ao;
  const float* theta = c.T.actor + (long long)agent * c.L.na_stride;
  const float* nr = c.T.norm + (long long)agent * c.L.norm_stride;
  const float* u = c.noise + (long long)agent * ((3LL * B + E) * A) + (long long)(2 * B + row) * A;
  const float ls_init = c.per_state_std ? logf(std_mult) - logf(kLog2) : logf(std_mult);
  const float floor_ls = logf(1e-3f);
  const int half = c.nmod == 2 ? E / 2 : E;
  const int net = row / half, il = row - net * half;
  float* dout = c.daOut + ((long long)agent * c.Rs + row) * Ao;
  for (int j = 0; j < A; ++j) {
    const GaussRow g = gauss_row(c, out, theta, j, ls_init, floor_ls);
    const float sd = expf(g.ls);
    float a = g.mean + sd * u[j];
    float cm = 1.f;
    if (clip) {          // tf.clip_by_value: the gradient passes on the closed interval
      const float lim = nr[c.L.off_act_limit + j];
      cm = (a >= -lim && a <= lim) ? 1.f : 0.f;
      a = fminf(fmaxf(a, -lim), lim);
    }
    const float inv_sd = 1.f / nstd(nr[c.L.off_m_a_std + j]);
    if (!bwd) {
      c.Xm[(((long long)agent * 2 + net) * E + il) * SA + S + j] = (a - nr[c.L.off_m_a_mean + j]) * inv_sd;
    } else {
      const float da = c.mdXa[(((long long)agent * 2 + net) * E + il) * A + j] * inv_sd * cm;
      dout[j] = da;
      const float dl = da * sd * u[j] * g.dls;
      if (c.per_state_std) dout[A + j] = dl;
      else c.dls[((long long)agent * c.Rs + row) * A + j] = dl;
    }
  }
}

// stats[agent*8 + 0] = MSE (sum over the models of their per-half means).  grid: (ceil(n/128)), block 128
__global__ void k_exp_stats(KCtx c, float* __restrict__ stats) {
  const int agent = blockIdx.x * blockDim.x + threadIdx.x;
  if (agent >= c.n_agents) return;
  stats[agent * 8] = c.mse_part[agent * 2] + (c.nmod == 2 ? c.mse_part[agent * 2 + 1] : 0.f);
}

// grad_final = (1 - eps) neg_pg + eps mse_grad (two rounded products, one rounded sum, like the reference's eager ops,
// trpo.py:150-158 / ppo.py:213), with the reference's logged norms: sums of per-tensor L2 norms (trpo.py:160-163).
// seg[0..nseg] = tensor boundaries in the flat layout.  stats[agent*8 + {6: norm_pg, 7: norm_MSE}].  grid: (n_agents), block 256
struct BlendSeg { int nseg; long long b[8]; };
__global__ void k_grad_blend(KCtx c, BlendSeg sg, const float* __restrict__ neg_pg, const float* __restrict__ mse_g,
                             const float* __restrict__ eps, float* __restrict__ out, float* __restrict__ stats) {
  __shared__ float sh[32];
  const int agent = blockIdx.x;
  const long long o = (long long)agent * c.L.na_stride;
  const float e = eps[agent], om = (float)(1.0 - (double)e);
  float npg = 0.f, nms = 0.f;
  for (int t = 0; t < sg.nseg; ++t) {
    float a2 = 0.f, b2 = 0.f;
    for (long long i = sg.b[t] + threadIdx.x; i < sg.b[t + 1]; i += blockDim.x) {
      const float a = neg_pg[o + i], b = mse_g[o + i];
      a2 += a * a; b2 += b * b;
      out[o + i] = __fadd_rn(__fmul_rn(om, a), __fmul_rn(e, b));
    }
    npg += sqrtf(block_sum(a2, sh)); nms += sqrtf(block_sum(b2, sh));
  }
  for (long long i = c.L.na + threadIdx.x; i < c.L.na_stride; i += blockDim.x) out[o + i] = 0.f;
  if (threadIdx.x == 0 && stats) { stats[agent * 8 + 6] = npg; stats[agent * 8 + 7] = nms; }
}

}  // namespace saceo
