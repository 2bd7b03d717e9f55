#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
namespace sdod {
int pack_rows(cudaStream_t s, const float* src, void* dst_bf16, int N, int K, int Kpad, const int* rowmap);
int pack_rows_window(cudaStream_t s, const float* src, void* dst_bf16, int N, int K, long long ld, int col0);
int add_f32(cudaStream_t s, const float* a, const float* b, float* dst, int n);
int gather_f32(cudaStream_t s, const float* src, float* dst, int n, const int* map);
int silu_f32_to_bf16(cudaStream_t s, const float* x, void* y_bf16, size_t n, int apply_silu);
int latent_prequant(cudaStream_t s, const float* z, void* y_bf16, size_t rows, const float* w, const float* b, float inv_scale);
int vae_post(cudaStream_t s, const float* x, uint8_t* u8, float* img, size_t n);
int broadcast_rows(cudaStream_t s, float* dst, const float* src, int rows, int width);
int fill_f32(cudaStream_t s, float* x, size_t n, float v);
int scale_f32(cudaStream_t s, float* x, size_t n, float v);
// out[r, :] = tok_emb[tokens[r], :] + pos_emb[r % T, :]   (CLIP text embeddings; ids clamped to the table)
int clip_embed(cudaStream_t s, const int* tokens, const float* tok_emb, const float* pos_emb, float* out, int rows, int T, int D, int vocab);
int cast_bf16_to_f32(cudaStream_t s, const void* x_bf16, float* y, size_t n);
}  // namespace sdod
