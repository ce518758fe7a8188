// mask_select.cuh -- which of the decoder's four masks a single-mask call returns.
//
// The SAM ONNX export (return_single_mask) scores the masks as  iou + (num_points - 2.5) * [1000, 0, 0, 0]  and takes the
// first maximum; dlimgedit always sends two prompt points (point + padding, or the two box corners), so mask 0 is
// penalised by 500 and the best of masks 1..3 wins unless its predicted IoU is absurdly low
// (reference: src/segmentation.cpp:95-118 feeds two points; the rule itself lives in the exported model).
// Used by select_masks_kernel (decoder_kernels.cu) and by the fused upscaling epilogue (gemm.cu), which must agree.
#pragma once

namespace dlimg {

__device__ __forceinline__ int best_mask_index(float const* iou4) {
    float best = iou4[0] + (2.0f - 2.5f) * 1000.0f;
    int bi = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        float const v = iou4[i] + (2.0f - 2.5f) * 0.0f;
        if (v > best) { best = v; bi = i; }
    }
    return bi;
}

// Which mask planes one decoder pass produces (Epilogue::fuse_mode of the fuse = 2 GEMM, SamModel::decode).
enum MaskMode : int { MASKS_ALL = 0, MASKS_MULTI = 1, MASKS_BEST = 2 };  // all four | masks 1..3 | the selected one

}  // namespace dlimg
