/* dlimg_b200.h -- C ABI of the B200-native dlimgedit segmentation engine (libdlimgedit.so).
 *
 * Part 1 re-declares, layout-identically, the dynamic-loading interface of Acly/dlimgedit so that this
 * library is a binary drop-in for the segmentation path: an application built against the reference's
 * own headers (dlimgedit.hpp -> detail/dlimgedit.h) links or dlopens this .so unchanged.
 *   reference interface: src/include/dlimgedit/detail/dlimgedit.h:24-70, filled in src/dlimgedit.cpp:102-117.
 *
 * Part 2 is additive (the reference ABI is strictly one image / one prompt per call, host buffers
 * only): batched, device-resident entry points behind a second exported symbol.  Nothing in part 1
 * changes size or order.
 *
 * No torch / C++ types cross this boundary: plain pointers, ints and sizes.
 */
#ifndef DLIMG_B200_H_
#define DLIMG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(DLIMG_B200_BUILD)
#    define DLIMG_B200_EXPORT __attribute__((visibility("default")))
#else
#    define DLIMG_B200_EXPORT
#endif

/* ------------------------------------------------------------------------------------------------
 * Part 1 -- drop-in boundary (same names as the reference so either header may be used)
 * ---------------------------------------------------------------------------------------------- */
#ifndef DLIMGEDIT_H_ /* the reference's own header already declares these */

typedef struct dlimg_Environment_* dlimg_Environment;   /* dlimgedit.h:24 */
typedef struct dlimg_Segmentation_* dlimg_Segmentation; /* dlimgedit.h:25 */

/* dlimgedit.h:27-33.  `channels` carries the dlimg::Channels enum VALUE (dlimgedit.hpp:29):
 * 1 mask, 3 rgb, 4 rgba, 5 bgra, 6 argb; bytes per pixel = channels > 4 ? 4 : channels.
 * `stride` is the row pitch in bytes (>= width * bytes per pixel). */
typedef struct dlimg_ImageView {
    int width;
    int height;
    int channels;
    int stride;
    uint8_t* pixels;
} dlimg_ImageView;

typedef enum dlimg_Backend { dlimg_cpu, dlimg_gpu } dlimg_Backend; /* dlimgedit.h:35 */

typedef struct dlimg_Options { /* dlimgedit.h:37-40 */
    dlimg_Backend backend;
    char const* model_directory;
} dlimg_Options;

typedef enum dlimg_Result { dlimg_success, dlimg_error } dlimg_Result; /* dlimgedit.h:42 */

/* dlimgedit.h:44-68 -- 13 slots, this order. */
struct dlimg_Api {
    /* dlimg_gpu: 1 iff device 0 (or $DLIMG_B200_DEVICE) has compute capability >= 10.0.
     * dlimg_cpu: always 0 -- this engine has no CPU path (reference: environment.cpp:120-122). */
    int (*is_backend_supported)(dlimg_Backend);

    /* reference dlimgedit.cpp:46-50.  model_directory must exist; weights are loaded lazily from
     * <model_directory>/segmentation/ on first use (environment.cpp:144-146). */
    dlimg_Result (*create_environment)(dlimg_Environment*, dlimg_Options const*);
    void (*destroy_environment)(dlimg_Environment);

    /* reference dlimgedit.cpp:52-59 -> SegmentationImpl::process (segmentation.cpp:121-129).
     * Host pixels are borrowed for the duration of the call.  *out is assigned before the work
     * starts so that on failure the caller still owns a destroyable handle. */
    dlimg_Result (*process_image_for_segmentation)(dlimg_Segmentation* out, dlimg_ImageView const*,
                                                   dlimg_Environment);
    /* reference dlimgedit.cpp:61-68 -> SegmentationImpl::compute_mask (segmentation.cpp:131-174).
     * point: 2 ints (x, y) or NULL; region: 4 ints (tl.x, tl.y, br.x, br.y) or NULL; exactly one is
     * non-NULL.  out_masks[3]: out_masks[1] == NULL selects the single-mask decoder, which writes
     * out_masks[0] only and leaves out_accuracys untouched; otherwise all three must be non-NULL and
     * receive mask tokens 1..3 plus their IoU predictions.  Masks are W*H bytes, packed, 0 or 255. */
    dlimg_Result (*get_segmentation_mask)(dlimg_Segmentation, int const* point, int const* region,
                                          uint8_t** out_masks, float* out_accuracys);
    void (*get_segmentation_extent)(dlimg_Segmentation, int* out_extent /* w, h */);
    void (*destroy_segmentation)(dlimg_Segmentation);

    /* reference dlimgedit.cpp:77-79 (BiRefNet).  Out of scope for this engine: always dlimg_error. */
    dlimg_Result (*segment_objects)(dlimg_ImageView const*, uint8_t* out_mask, dlimg_Environment);

    /* reference dlimgedit.cpp:81-90 (stb image I/O).  load_image: PNG (every colour type and bit depth, palettes, tRNS, Adam7; 16-bit samples keep the high byte),
     * JPEG (8-bit, Huffman, sequential or progressive; stb_image's IDCT / chroma upsampling / colour conversion
     * restated), BMP (palettes, 16 / 24 / 32 bits, bit fields; no RLE), TGA (types 1 / 2 / 3 and their run-length
     * forms) and binary PGM / PPM; arithmetic-coded / lossless JPEG are refused.  save_image: PNG (filtered
     * scanlines, fixed-Huffman deflate) for mask / rgb / rgba.  Pixel buffers of load_image / create_image are released
     * with destroy_image; while a GPU environment is alive they are page-locked and recycled by size, so process /
     * get_segmentation_mask on them copy over PCIe without pageable staging (csrc/image_pool.hpp). */
    dlimg_Result (*load_image)(char const*, int* out_extent, int* out_channels, uint8_t** out_pixels);
    dlimg_Result (*save_image)(dlimg_ImageView const*, char const*);
    uint8_t* (*create_image)(int w, int h, int channels);
    void (*destroy_image)(uint8_t const*);

    char const* (*last_error)(void); /* thread-local in this engine (global in the reference) */
};

#endif /* DLIMGEDIT_H_ */

/* The one symbol a drop-in must export (dlimgedit.h:70). */
DLIMG_B200_EXPORT struct dlimg_Api const* dlimg_init(void);

/* ------------------------------------------------------------------------------------------------
 * Part 2 -- additive batch / device-resident extension
 * ---------------------------------------------------------------------------------------------- */

typedef struct dlimg_b200_Prompt {
    int kind;     /* 0 = point (x0, y0), 1 = region (x0, y0) top-left, (x1, y1) bottom-right */
    int x0, y0, x1, y1;
} dlimg_b200_Prompt;

/* One row of the optional per-kernel profile (see profile_enable / profile_read). */
typedef struct dlimg_b200_ProfileEntry {
    char name[32];     /* kernel category, e.g. "gemm_tcgen05_f16" */
    uint64_t launches;
    double ms;         /* summed CUDA-event time of those launches */
    double flops;      /* algorithmic floating point operations (2*M*N*K for GEMMs), 0 if not tracked */
    double bytes;      /* algorithmic bytes moved, 0 if not tracked */
} dlimg_b200_ProfileEntry;

typedef struct dlimg_b200_Stats {
    uint64_t kernel_launches; /* kernels of THIS library launched since the environment was created */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
} dlimg_b200_Stats;

struct dlimg_b200_Ext {
    uint32_t struct_size; /* sizeof(struct dlimg_b200_Ext) of the library, for forward compatibility */
    uint32_t abi_version;

    /* Work submitted through the *_batch calls below goes to `cuda_stream` (a cudaStream_t; NULL =
     * the environment's own stream).  Batch calls are asynchronous with respect to the host unless
     * stated otherwise; use `synchronize` or your own events on that stream.  Switching streams needs no
     * synchronisation by the caller: everything already queued on the old stream is ordered in front of
     * whatever is submitted to the new one (event dependency), because workspaces and embeddings are shared.
     * DEVICE buffers passed to any call below are read and written on that stream: the environment's own stream is
     * non-blocking, so a caller that fills them on another stream shares it here or synchronises first. */
    dlimg_Result (*set_stream)(dlimg_Environment, void* cuda_stream);
    dlimg_Result (*synchronize)(dlimg_Environment); /* the work stream and the library's two copy streams */
    dlimg_Result (*get_stats)(dlimg_Environment, dlimg_b200_Stats*);

    /* Encode `count` images.  views[i].pixels is a HOST pointer when pixels_on_device == 0 or a DEVICE
     * pointer otherwise.  Host pixels are copied straight from the caller's memory on the library's upload
     * stream (one copy per run of packed images that are contiguous in memory; truly asynchronous only from
     * page-locked memory) and are free again when the call returns; the encoder itself keeps running.  All
     * images of one call must share width, height, channels; stride may differ.  out[i] receives a new
     * segmentation handle (destroy with dlimg_Api.destroy_segmentation). */
    dlimg_Result (*process_batch)(dlimg_Environment, dlimg_ImageView const* views, int count,
                                  int pixels_on_device, dlimg_Segmentation* out);

    /* Answer `count` prompts; prompt i refers to segs[i] (handles may repeat, and may belong to different
     * images of any extent: up to $DLIMG_B200_MAX_PROMPTS prompts share one decoder pass).  multi == 0: one mask
     * per prompt (best of tokens 1..3 by predicted IoU, as the single-mask decoder graph does);
     * multi == 1: three masks per prompt (tokens 1..3).  masks_out[i] points at n*W_i*H_i bytes
     * (n = multi ? 3 : 1) on the device (masks_on_device == 1) or host (0: complete when the call returns;
     * 2: asynchronous -- the downloads run on the library's copy-out stream under the decoder of the following
     * calls and the buffers, ideally page-locked, are complete after `synchronize`).  ious_out: count*n floats
     * (same memory space as masks), may be NULL. */
    dlimg_Result (*compute_masks_batch)(dlimg_Environment, dlimg_Segmentation const* segs,
                                        dlimg_b200_Prompt const* prompts, int count, int multi,
                                        uint8_t* const* masks_out, float* ious_out, int masks_on_device);

    /* Copies the (1, 256, 64, 64) float32 image embedding (NCHW, the layout of the reference's
     * `image_embeddings` tensor, segmentation.cpp:124) to a host buffer of 256*64*64 floats. */
    dlimg_Result (*get_embedding)(dlimg_Segmentation, float* out_host);
    /* Low-resolution mask logits (4, 256, 256) float32 and the 4 IoU predictions for one prompt: the
     * decoder graph's `low_res_masks` / `iou_predictions` before selection (segmentation.cpp:24). */
    dlimg_Result (*get_low_res_logits)(dlimg_Segmentation, dlimg_b200_Prompt const*, float* out_logits_host,
                                       float* out_iou_host);

    /* Stand-alone pre/post-processing stages, device pointers, on the environment's stream:
     *  - resize_longest_side: reference ResizeLongestSide::resize + dlimg::resize (segmentation.cpp:60-70,
     *    image.cpp:37-51).  Writes packed u8 (out_h, out_w, bpp); returns the extent via out_extent.
     *  - image_tensor: reference create_image_tensor (segmentation.cpp:81-106), float32 (h, w, 3), 0..255.
     *  - mask_postprocess: upsample 256->1024, crop, resize to (w, h), threshold > 0 -> 0/255
     *    (decoder graph post-processing + write_mask_image, segmentation.cpp:108-116).
     *  - threshold_mask: write_mask_image alone on a (th, tw) float32 logits plane. */
    dlimg_Result (*resize_longest_side)(dlimg_Environment, dlimg_ImageView const* dev_view, int max_side,
                                        uint8_t* dev_out, int* out_extent);
    dlimg_Result (*image_tensor)(dlimg_Environment, dlimg_ImageView const* dev_view, float* dev_out);
    dlimg_Result (*mask_postprocess)(dlimg_Environment, float const* dev_low_res, int count, int w, int h,
                                     uint8_t* dev_out);
    dlimg_Result (*threshold_mask)(dlimg_Environment, float const* dev_logits, int th, int tw, int w, int h,
                                   uint8_t* dev_out);

    /* Per-kernel timing with CUDA events on the launching stream (adds two event records per launch, so
     * switch it on only for an attribution pass).  profile_read waits for the recorded events, writes up to
     * `capacity` non-empty categories, clears the recording and returns the number written through *count. */
    dlimg_Result (*profile_enable)(dlimg_Environment, int on);
    dlimg_Result (*profile_read)(dlimg_Environment, dlimg_b200_ProfileEntry* out, int capacity, int* count);

    /* abi_version >= 2.  get_embedding without the wait: the NCHW copy is queued behind the encoder and the
     * device->host transfer runs on the library's copy-out stream, overlapping later work (e.g. the next
     * process_batch, whose host->device upload runs on a third stream).  out_host should be page-locked and must
     * stay valid until `synchronize` returns. */
    dlimg_Result (*get_embedding_async)(dlimg_Segmentation, float* out_host);

    /* abi_version >= 3.  The same, as IEEE half precision: 256*64*64 uint16 values, half the bytes over PCIe (the
     * embedding is converted on the device on its way out).  For callers that keep embeddings on the host and are bound
     * by the download -- eight ranks on one host share its copy bandwidth. */
    dlimg_Result (*get_embedding_f16_async)(dlimg_Segmentation, uint16_t* out_host);
};

DLIMG_B200_EXPORT struct dlimg_b200_Ext const* dlimg_b200_ext_init(void);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* DLIMG_B200_H_ */
