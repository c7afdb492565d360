// loss_adam.cu — K9 (fused masked dot + BCE-with-logits + mean, forward and backward) and K14 (flat Adam).
//
// K9 restates Model.forward's loss (Downstream/Text/model/model.py:53-68) and ModelCPC.forward's
// (model.py:120-133): with emb = encoder output viewed [B, S+1, 2, D] (slot [:, :, 0] = history item,
// [:, :, 1] = sampled negative) and prec = user-encoder output [B, S, D]:
//   p[b,t] = <prec[b,t], emb[b,t+1,0]>      n[b,t] = <prec[b,t], emb[b,t,1]>
//   loss = mean_V softplus(-p) + mean_V softplus(n),   V = {(b,t): log_mask[b,t] != 0}   (CPC: V = {(b,S-1)})
// The reference reaches V through torch.where (a device->host sync); here the mean is a deterministic
// two-stage reduction on the device and nothing synchronises.
#include "a4r_common.cuh"

namespace {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_BLOCKS = 296;

A4R_DEVICE float softplus_stable(float x) { return fmaxf(x, 0.0f) + log1pf(__expf(-fabsf(x))); }
A4R_DEVICE float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

A4R_DEVICE bool pos_valid(const a4r_bce_args& a, int64_t b, int t) {
  if (a.cpc) return t == a.S - 1;
  return a.log_mask[b * a.S + t] != 0.0f;
}

__global__ void __launch_bounds__(LOSS_THREADS) bce_fwd_kernel(const a4r_bce_args a, float* __restrict__ partial) {
  __shared__ float red[LOSS_THREADS / 32][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = static_cast<int>(a.D), S = static_cast<int>(a.S);
  const __nv_bfloat16* prec = static_cast<const __nv_bfloat16*>(a.prec);
  const __nv_bfloat16* emb = static_cast<const __nv_bfloat16*>(a.emb);
  float sp = 0.0f, sn = 0.0f, cnt = 0.0f;
  const int64_t total = a.B * S;
  for (int64_t w = static_cast<int64_t>(blockIdx.x) * (LOSS_THREADS / 32) + warp; w < total;
       w += static_cast<int64_t>(gridDim.x) * (LOSS_THREADS / 32)) {
    const int64_t b = w / S;
    const int t = static_cast<int>(w % S);
    const __nv_bfloat16* h = prec + w * D;
    const __nv_bfloat16* ep = emb + ((b * (S + 1) + t + 1) * 2 + 0) * D;
    const __nv_bfloat16* en = emb + ((b * (S + 1) + t) * 2 + 1) * D;
    float p = 0.0f, n = 0.0f;
    for (int c = lane * 8; c < D; c += 256) {
      const uint4 hv = *reinterpret_cast<const uint4*>(h + c);
      const uint4 pv = *reinterpret_cast<const uint4*>(ep + c);
      const uint4 nv = *reinterpret_cast<const uint4*>(en + c);
      const uint32_t hh[4] = {hv.x, hv.y, hv.z, hv.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w}, nn[4] = {nv.x, nv.y, nv.z, nv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack_bf16x2(hh[e]), y = unpack_bf16x2(pp[e]), z = unpack_bf16x2(nn[e]);
        p += x.x * y.x + x.y * y.y;
        n += x.x * z.x + x.y * z.y;
      }
    }
    p = warp_sum(p);
    n = warp_sum(n);
    if (lane == 0) {
      a.pos_score[w] = p;
      a.neg_score[w] = n;
      if (pos_valid(a, b, t)) {
        sp += softplus_stable(-p);
        sn += softplus_stable(n);
        cnt += 1.0f;
      }
    }
  }
  if (lane == 0) {
    red[warp][0] = sp;
    red[warp][1] = sn;
    red[warp][2] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float acc = 0.0f;
    for (int i = 0; i < LOSS_THREADS / 32; ++i) acc += red[i][threadIdx.x];
    partial[blockIdx.x * 3 + threadIdx.x] = acc;
  }
}

__global__ void bce_finalize_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ loss,
                                    float* __restrict__ count) {
  if (threadIdx.x == 0) {
    float sp = 0.0f, sn = 0.0f, cnt = 0.0f;
    for (int i = 0; i < nblocks; ++i) {
      sp += partial[i * 3 + 0];
      sn += partial[i * 3 + 1];
      cnt += partial[i * 3 + 2];
    }
    *loss = sp / cnt + sn / cnt;  // 0/0 = NaN when no position is valid, as the reference's mean over an empty set
    *count = cnt;
  }
}

// one warp per (b,t): writes d_prec[b,t], d_emb[b,t+1,0] (target role of the next item) and d_emb[b,t,1]
__global__ void __launch_bounds__(LOSS_THREADS) bce_bwd_kernel(const a4r_bce_args a, const float* __restrict__ grad_out,
                                                              __nv_bfloat16* __restrict__ d_prec,
                                                              __nv_bfloat16* __restrict__ d_emb) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = static_cast<int>(a.D), S = static_cast<int>(a.S);
  const __nv_bfloat16* prec = static_cast<const __nv_bfloat16*>(a.prec);
  const __nv_bfloat16* emb = static_cast<const __nv_bfloat16*>(a.emb);
  const float go = (grad_out != nullptr ? *grad_out : 1.0f) / *a.count;
  const int64_t total = a.B * S;
  for (int64_t w = static_cast<int64_t>(blockIdx.x) * (LOSS_THREADS / 32) + warp; w < total;
       w += static_cast<int64_t>(gridDim.x) * (LOSS_THREADS / 32)) {
    const int64_t b = w / S;
    const int t = static_cast<int>(w % S);
    const bool valid = pos_valid(a, b, t);
    const float gp = valid ? -sigmoidf_(-a.pos_score[w]) * go : 0.0f;  // d loss / d p
    const float gn = valid ? sigmoidf_(a.neg_score[w]) * go : 0.0f;    // d loss / d n
    const int64_t ip = ((b * (S + 1) + t + 1) * 2 + 0) * D, in = ((b * (S + 1) + t) * 2 + 1) * D;
    for (int c = lane * 8; c < D; c += 256) {
      const uint4 hv = *reinterpret_cast<const uint4*>(prec + w * D + c);
      const uint4 pv = *reinterpret_cast<const uint4*>(emb + ip + c);
      const uint4 nv = *reinterpret_cast<const uint4*>(emb + in + c);
      const uint32_t hh[4] = {hv.x, hv.y, hv.z, hv.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w}, nn[4] = {nv.x, nv.y, nv.z, nv.w};
      uint32_t oh[4], op[4], on[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack_bf16x2(hh[e]), y = unpack_bf16x2(pp[e]), z = unpack_bf16x2(nn[e]);
        oh[e] = pack_bf16x2(gp * y.x + gn * z.x, gp * y.y + gn * z.y);
        op[e] = pack_bf16x2(gp * x.x, gp * x.y);
        on[e] = pack_bf16x2(gn * x.x, gn * x.y);
      }
      *reinterpret_cast<uint4*>(d_prec + w * D + c) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
      *reinterpret_cast<uint4*>(d_emb + ip + c) = make_uint4(op[0], op[1], op[2], op[3]);
      *reinterpret_cast<uint4*>(d_emb + in + c) = make_uint4(on[0], on[1], on[2], on[3]);
      // the two slots no (b,t) pair targets: history slot 0 and the negative slot S
      if (t == 0) *reinterpret_cast<uint4*>(d_emb + ((b * (S + 1)) * 2 + 0) * D + c) = make_uint4(0, 0, 0, 0);
      if (t == S - 1) *reinterpret_cast<uint4*>(d_emb + ((b * (S + 1) + S) * 2 + 1) * D + c) = make_uint4(0, 0, 0, 0);
    }
  }
}

// torch.optim.Adam (no amsgrad, no weight decay unless given):  one flat fp32 segment per launch
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float beta1, float beta2, float eps,
                            float weight_decay, float bc1, float bc2_sqrt, float grad_scale,
                            const float* __restrict__ bias_corr) {
  if (bias_corr != nullptr) {   // step recorded in a CUDA graph: the step-dependent factors come from device memory
    bc1 = bias_corr[0];
    bc2_sqrt = bias_corr[1];
  }
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float grad = g[i] * grad_scale;
    const float w = p[i];
    if (weight_decay != 0.0f) grad += weight_decay * w;
    const float mi = beta1 * m[i] + (1.0f - beta1) * grad;
    const float vi = beta2 * v[i] + (1.0f - beta2) * grad * grad;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = w - (lr / bc1) * (mi / denom);
  }
}

int check_bce(const a4r_bce_args* a) {
  A4R_CHECK_ARG(a != nullptr, "bce: args is NULL");
  A4R_CHECK_ARG(a->prec && a->emb && a->pos_score && a->neg_score && a->loss && a->count, "bce: NULL pointer");
  A4R_CHECK_ARG(a->cpc || a->log_mask, "bce: log_mask is required unless cpc");
  A4R_CHECK_ARG(a->B >= 1 && a->S >= 1 && a->D >= 8 && a->D % 8 == 0, "bce: bad B/S/D");
  A4R_CHECK_ARG(a4r_aligned16(a->prec) && a4r_aligned16(a->emb), "bce: pointers must be 16B aligned");
  return a4r_device_check();
}

}  // namespace

extern "C" size_t a4r_bce_workspace_bytes(void) { return LOSS_MAX_BLOCKS * 3 * sizeof(float); }

extern "C" int a4r_bce_loss_fwd(const a4r_bce_args* a, void* workspace, size_t workspace_bytes, a4r_stream_t stream_) {
  int rc = check_bce(a);
  if (rc != A4R_OK) return rc;
  if (workspace == nullptr || workspace_bytes < a4r_bce_workspace_bytes())
    return a4r_set_error(A4R_EWORKSPACE, "bce: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t total = a->B * a->S;
  int64_t blocks = (total + LOSS_THREADS / 32 - 1) / (LOSS_THREADS / 32);
  if (blocks > LOSS_MAX_BLOCKS) blocks = LOSS_MAX_BLOCKS;
  bce_fwd_kernel<<<static_cast<int>(blocks), LOSS_THREADS, 0, stream>>>(*a, static_cast<float*>(workspace));
  A4R_LAUNCH_OK();
  bce_finalize_kernel<<<1, 32, 0, stream>>>(static_cast<const float*>(workspace), static_cast<int>(blocks), a->loss,
                                            a->count);
  A4R_LAUNCH_OK();
  a4r_count_launch(2);
  return A4R_OK;
}

extern "C" int a4r_bce_loss_bwd(const a4r_bce_args* a, const float* grad_out, void* d_prec, void* d_emb,
                                a4r_stream_t stream_) {
  int rc = check_bce(a);
  if (rc != A4R_OK) return rc;
  A4R_CHECK_ARG(d_prec && d_emb && a4r_aligned16(d_prec) && a4r_aligned16(d_emb), "bce bwd: bad outputs");
  const int64_t total = a->B * a->S;
  int64_t blocks = (total + LOSS_THREADS / 32 - 1) / (LOSS_THREADS / 32);
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  bce_bwd_kernel<<<static_cast<int>(blocks), LOSS_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
      *a, grad_out, static_cast<__nv_bfloat16*>(d_prec), static_cast<__nv_bfloat16*>(d_emb));
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}

extern "C" int a4r_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                             a4r_stream_t stream_) {
  A4R_CHECK_ARG(p && g && m && v, "adam: NULL pointer");
  A4R_CHECK_ARG(n >= 0 && step >= 1, "adam: bad n/step");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (n == 0) return A4R_OK;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, static_cast<float>(bc1), static_cast<float>(sqrt(bc2)),
      grad_scale, nullptr);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}

extern "C" int a4r_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, const float* bias_corr, float grad_scale,
                                 a4r_stream_t stream_) {
  A4R_CHECK_ARG(p && g && m && v && bias_corr, "adam: NULL pointer");
  A4R_CHECK_ARG(n >= 0, "adam: bad n");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (n == 0) return A4R_OK;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, 1.0f, 1.0f, grad_scale, bias_corr);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}
