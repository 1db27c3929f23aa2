// Host entry points of the DeiT-Tiny trunk (implemented in encoder.cu, exported through api.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

int64_t rvk_encoder_weight_bytes_impl(int training);
int64_t rvk_encoder_workspace_bytes_impl(int batch, int training, int chunk_images);
int rvk_encoder_prepare_weights_impl(const void* const* params, void* wbuf, int training, cudaStream_t s);
int rvk_encoder_forward_impl(const void* const* params, const void* wbuf, const void* images, int image_fmt, const float* norm6_host, int batch, int training,
                             int chunk_images, void* workspace, float* features, cudaStream_t s);
void rvk_set_side_stream_impl(int on);
int64_t rvk_encoder_saved_offset_impl(int batch, int block, int which);
int rvk_encoder_backward_impl(const void* const* params, const void* wbuf, void* workspace, const float* dfeatures,
                              int batch, int chunk_images, void* const* grads, int stage_begin, int stage_end, cudaStream_t s);
