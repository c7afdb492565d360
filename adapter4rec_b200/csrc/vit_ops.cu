// vit_ops.cu — K15 support kernels for the ViT item encoder (HBM-bound):
//   a4r_patchify       [N,C,R,R] f32 image -> [N*P, C*ps*ps] bf16 patch rows.  ViTPatchEmbeddings is a ps x ps conv with
//                      stride ps (non-overlapping), i.e. exactly the GEMM  patches · W[out, C*ps*ps]ᵀ + b  that
//                      gemm_sm100.cu runs; this kernel is the im2col, which for non-overlapping patches is a pure
//                      permutation fused with the f32 -> bf16 conversion.
//   a4r_vit_assemble   token assembly of ViTEmbeddings.forward / SoftPrompt.forward (Downstream/CV/model/model.py:523-535):
//                      out[n,0] = cls + pos[0]; out[n,1+p] = patch_emb[n,p] + pos[1+p]; out[n,1+P+t] = prompt[t]
//                      (prompt tokens are appended AFTER the position embeddings are added, as the reference does).
#include "a4r_common.cuh"

namespace {

// one thread per 8 consecutive output elements (16 B store): output column k = (c*ps + py)*ps + px, 8 | ps
__global__ void patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int64_t N, int C, int R,
                                int ps) {
  const int G = R / ps;                 // patches per side
  const int K = C * ps * ps;
  const int64_t total = N * G * G * (K / 8);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int kc = static_cast<int>(i % (K / 8));
    const int64_t row = i / (K / 8);
    const int64_t n = row / (G * G);
    const int pidx = static_cast<int>(row % (G * G));
    const int gy = pidx / G, gx = pidx % G;
    const int k = kc * 8;
    const int c = k / (ps * ps), rem = k % (ps * ps);
    const int py = rem / ps, px = rem % ps;  // px..px+7 stay inside the patch row because 8 | ps
    const float* src = img + ((n * C + c) * R + (gy * ps + py)) * static_cast<int64_t>(R) + gx * ps + px;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + 4));
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    st_na_v4(out + row * K + k, o);
  }
}

__global__ void vit_assemble_kernel(const __nv_bfloat16* __restrict__ patch_emb, const __nv_bfloat16* __restrict__ cls,
                                    const __nv_bfloat16* __restrict__ pos, const __nv_bfloat16* __restrict__ prompt,
                                    __nv_bfloat16* __restrict__ out, int64_t N, int P, int T, int H) {
  const int Ltot = 1 + P + T;
  const int chunks = H / 8;
  const int64_t total = N * Ltot * chunks;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % chunks);
    const int64_t tok = i / chunks;
    const int64_t n = tok / Ltot;
    const int t = static_cast<int>(tok % Ltot);
    uint4 o;
    if (t > P) {
      o = __ldg(reinterpret_cast<const uint4*>(prompt + static_cast<int64_t>(t - 1 - P) * H + ch * 8));
    } else {
      const uint4 a = t == 0 ? __ldg(reinterpret_cast<const uint4*>(cls + ch * 8))
                             : ld_nc_v4(patch_emb + (n * P + (t - 1)) * H + ch * 8);
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(pos + static_cast<int64_t>(t) * H + ch * 8));
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
      uint32_t ow[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(bw[e]);
        ow[e] = pack_bf16x2(x.x + y.x, x.y + y.y);
      }
      o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    st_na_v4(out + tok * H + ch * 8, o);
  }
}

}  // namespace

extern "C" int a4r_patchify(const float* images, void* out, int64_t N, int64_t C, int64_t R, int64_t ps,
                            a4r_stream_t stream_) {
  A4R_CHECK_ARG(images && out, "patchify: NULL pointer");
  A4R_CHECK_ARG(N >= 0 && C >= 1 && ps >= 8 && ps % 8 == 0 && R >= ps && R % ps == 0,
                "patchify: need ps %% 8 == 0 and R %% ps == 0 (C=%lld R=%lld ps=%lld)", (long long)C, (long long)R, (long long)ps);
  A4R_CHECK_ARG(a4r_aligned16(images) && a4r_aligned16(out), "patchify: pointers must be 16B aligned");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (N == 0) return A4R_OK;
  const int64_t total = N * (R / ps) * (R / ps) * (C * ps * ps / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  patchify_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      images, static_cast<__nv_bfloat16*>(out), N, static_cast<int>(C), static_cast<int>(R), static_cast<int>(ps));
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}

extern "C" int a4r_vit_assemble(const void* patch_emb, const void* cls, const void* pos, const void* prompt, void* out,
                                int64_t N, int64_t P, int64_t T, int64_t H, a4r_stream_t stream_) {
  A4R_CHECK_ARG(patch_emb && cls && pos && out, "vit_assemble: NULL pointer");
  A4R_CHECK_ARG(N >= 0 && P >= 1 && T >= 0 && H >= 8 && H % 8 == 0, "vit_assemble: bad N/P/T/H");
  A4R_CHECK_ARG(T == 0 || prompt != nullptr, "vit_assemble: prompt is required when T > 0");
  A4R_CHECK_ARG(a4r_aligned16(patch_emb) && a4r_aligned16(cls) && a4r_aligned16(pos) && a4r_aligned16(prompt) &&
                    a4r_aligned16(out), "vit_assemble: pointers must be 16B aligned");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (N == 0) return A4R_OK;
  const int64_t total = N * (1 + P + T) * (H / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  vit_assemble_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(patch_emb), static_cast<const __nv_bfloat16*>(cls),
      static_cast<const __nv_bfloat16*>(pos), static_cast<const __nv_bfloat16*>(prompt),
      static_cast<__nv_bfloat16*>(out), N, static_cast<int>(P), static_cast<int>(T), static_cast<int>(H));
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}
