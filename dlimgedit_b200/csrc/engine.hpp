// engine.hpp -- host-side objects behind the C ABI: EnvironmentImpl (device context, weights, workspaces)
// and SegmentationImpl (one encoded image).  They take the place of the reference's EnvironmentImpl
// (environment.hpp:20-42), Session (session.hpp:32-57) and SegmentationImpl (segmentation.hpp:47-62).
#pragma once

#include "../../include/dlimg_b200.h"
#include "common.hpp"
#include "kernels/prepost_kernels.cuh"
#include "model.hpp"
#include "profiler.hpp"

#include <map>
#include <tuple>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace dlimg {

class SegmentationImpl;

struct DeviceAxisPlan {
    DeviceBuffer<int> first;
    DeviceBuffer<float> weights;
    std::vector<int> first_host;  // the tile planner of the resize kernel reads it
    int taps = 0;
};

// Device memory from the stream-ordered allocator (cudaMallocAsync): per-call embedding stores are recycled from
// the driver's pool instead of paying a blocking cudaMalloc / cudaFree per batch.
class StreamBuffer {
  public:
    StreamBuffer(size_t bytes, cudaStream_t stream);
    ~StreamBuffer();
    StreamBuffer(StreamBuffer const&) = delete;
    StreamBuffer& operator=(StreamBuffer const&) = delete;
    uint8_t* bytes() const { return static_cast<uint8_t*>(ptr_); }
    // Cross-stream lifetime of an embedding store: `ready` is recorded on the work stream once the encoder has filled
    // the buffer (every consumer on another stream waits for it), `last_read` on the copy-out stream after each download,
    // `last_use` on the work stream after each decoder pass that read it (the work stream may have been switched with
    // set_stream since the allocation); the destructor makes the stream-ordered free wait for both.
    cudaEvent_t ready() const { return ready_; }
    void mark_read(cudaStream_t copy_stream);
    void mark_used(cudaStream_t work_stream);
    // Embedding stores only (guarded by the environment's mutex): number of images in the chunk, and the state of the 16-bit
    // copy of their embeddings that sits behind them -- converted once for the whole chunk by the first
    // get_embedding_f16_async of any of its images, `f16_ready` recorded behind that conversion.
    int images = 0;
    bool f16_done = false;
    cudaEvent_t f16_ready = nullptr;

  private:
    void* ptr_ = nullptr;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ready_ = nullptr, last_read_ = nullptr, last_use_ = nullptr;
    bool read_ = false, used_ = false;
    std::mutex mutex_;
};

// Page-locked host arena for small parameter uploads (prompt coordinates, descriptors, pointer tables).
class PinnedArena {
  public:
    explicit PinnedArena(size_t bytes);
    ~PinnedArena();
    // Returns `bytes` of pinned memory; when the arena wraps it first waits for `stream` so that earlier
    // asynchronous copies out of the arena have completed.
    void* take(size_t bytes, cudaStream_t stream);

  private:
    uint8_t* base_ = nullptr;
    size_t cap_ = 0, off_ = 0;
};

class EnvironmentImpl {
  public:
    explicit EnvironmentImpl(dlimg_Options const& opts);
    ~EnvironmentImpl();

    static bool is_supported(dlimg_Backend backend);
    static int device_ordinal();  // $DLIMG_B200_DEVICE or 0

    SamModel& model();  // loaded on first use (reference lazy.hpp:10-13)
    cudaStream_t stream() const { return user_stream_ ? user_stream_ : own_stream_; }
    // Switches the work stream.  Everything already queued on the old stream is ordered in front of whatever is
    // submitted to the new one (an event dependency, no host wait): the workspaces and embeddings are shared.
    void set_stream(cudaStream_t s);
    void synchronize();  // the work stream and both copy streams
    void bind_device() const;

    // Encodes `count` images (host or device pixels) and fills `out` with new handles.
    void process_batch(dlimg_ImageView const* views, int count, bool on_device, SegmentationImpl** out);
    // Prompts -> masks.  See dlimg_b200_Ext::compute_masks_batch.
    void compute_masks_batch(SegmentationImpl* const* segs, dlimg_b200_Prompt const* prompts, int count, bool multi,
                             uint8_t* const* masks_out, float* ious_out, int placement /* 0 host, 1 device, 2 host asynchronous */);
    void low_res_logits(SegmentationImpl& seg, dlimg_b200_Prompt const& prompt, float* logits_host, float* iou_host);

    // Stand-alone stages (device pointers)
    void resize_longest_side(dlimg_ImageView const& dev_view, int max_side, uint8_t* dev_out, int* out_extent);
    void image_tensor(dlimg_ImageView const& dev_view, float* dev_out);
    void mask_postprocess(float const* dev_low_res, int count, int w, int h, uint8_t* dev_out);
    void threshold_mask(float const* dev_logits, int th, int tw, int w, int h, uint8_t* dev_out);

    // debug: encode one device-resident image and copy a named activation (see model.cu tap names)
    size_t encode_tap(dlimg_ImageView const* dev_views, int count, char const* tap_name, float* dev_out, size_t capacity);

    void profile_enable(bool on);
    std::vector<Profiler::Total> profile_collect();
    void stats(dlimg_b200_Stats& out) const;

    std::mutex& mutex() { return mutex_; }
    int max_batch() const { return max_batch_; }
    int max_prompts() const { return max_prompts_; }

  private:
    friend class SegmentationImpl;
    // Binds this environment's counters and profiler to the calling thread for the duration of a call (the kernel
    // launch wrappers report to whichever environment is bound), and selects its device.
    struct Scope {
        explicit Scope(EnvironmentImpl const& env);
        ~Scope();
        EnvCounters* prev_counters;
        Profiler* prev_profiler;
    };
    EncoderWorkspace& encoder_ws(int batch);   // grown on demand up to max_batch_ images
    DecoderWorkspace& decoder_ws(int prompts); // grown on demand up to max_prompts_ prompts
    prepost::ResizeDeviceTables resize_tables(int in_w, int in_h, int out_w, int out_h);
    DeviceAxisPlan const& axis_plan(int in_size, int out_size);
    struct ChunkOut { float* emb_nchw; act_t* keys0; act_t* kvq0; };
    void encode_chunk(enc::ImageDesc const* host_descs, int batch, prepost::LongestSide const& size, int channels, ChunkOut const& out,
                      Tap* tap);
    void decode_chunk(SegmentationImpl* const* segs, dlimg_b200_Prompt const* prompts, int P, bool eager, int mask_mode);
    void prepare_input(dlimg_ImageView const& view, uint8_t const* dev_pixels, int dev_stride, prepost::LongestSide const& size,
                       int slot, enc::ImageDesc& desc);
    void release_graphs();

    int device_ = 0;
    int num_sms_ = 148;
    std::string model_dir_;
    std::mutex mutex_;       // serialises GPU submission per environment (Environment is thread-safe)
    std::once_flag model_once_;
    std::unique_ptr<SamModel> model_;
    cudaStream_t own_stream_ = nullptr;
    cudaStream_t user_stream_ = nullptr;
    cudaEvent_t stream_switch_ = nullptr;
    int max_batch_ = 32;
    int max_prompts_ = 64;
    mutable EnvCounters counters_;
    mutable Profiler profiler_;
    std::unique_ptr<EncoderWorkspace> enc_ws_;
    std::unique_ptr<DecoderWorkspace> dec_ws_;
    std::unique_ptr<PinnedArena> pinned_;
    // Host pixels are uploaded on a dedicated copy stream into one of two slot sets, so the upload of call i+1
    // overlaps the encoder of call i; embeddings leave on a second copy stream (get_embedding_async).
    cudaStream_t copy_in_ = nullptr, copy_out_ = nullptr;
    cudaEvent_t h2d_done_ = nullptr;
    cudaEvent_t input_free_[2] = {nullptr, nullptr};  // the encoder that read slot set k has finished
    bool input_used_[2] = {false, false};
    int input_flip_ = 0;
    DeviceBuffer<uint8_t> input_px_[2];   // uploaded originals, slots of input_slot_bytes_
    size_t input_slots_ = 0;              // slots per set
    DeviceBuffer<uint8_t> resized_px_;    // resized images (<= 1024 x 1024 x 4 each)
    size_t resized_slots_ = 0;
    DeviceBuffer<float> resize_scratch_;  // horizontal-pass intermediate of the two-pass resize fallback only
    size_t input_slot_bytes_ = 0;
    DeviceBuffer<enc::ImageDesc> descs_;
    DeviceBuffer<float> srgb_decode_, srgb_threshold_;
    std::map<std::pair<int, int>, DeviceAxisPlan> plans_;
    // device staging for host-destined masks, double-buffered: the download of one prompt group (copy-out stream)
    // overlaps the decoder of the next
    DeviceBuffer<uint8_t> mask_out_[2];
    cudaEvent_t mask_ready_ = nullptr, mask_free_[2] = {nullptr, nullptr};
    bool mask_used_[2] = {false, false};
    int host_group_ = 0;  // alternates the two device-side mask staging buffers across prompt groups AND calls
    DeviceBuffer<uint8_t*> plane_ptrs_;
    bool use_graphs_ = true;  // $DLIMG_B200_GRAPHS=0 forces eager launches
    struct Graph { cudaGraphExec_t exec = nullptr; uint64_t kernels = 0; };
    std::map<std::tuple<int, int, int, int>, Graph> encode_graphs_;  // (batch, w, h, channels)
    std::map<int, Graph> decode_graphs_;                             // prompts per pass
    template <typename F> Graph capture(cudaStream_t s, F const& body);
};

class SegmentationImpl {
  public:
    explicit SegmentationImpl(EnvironmentImpl& env) : env_(env) {}

    void process(dlimg_ImageView const& view);  // reference segmentation.cpp:121-129
    void compute_mask(int const* point, int const* region, uint8_t** out_masks, float* out_accuracy);  // :131-174
    void embedding_nchw(float* out_host);        // blocking
    void embedding_nchw_async(float* out_host);  // returns at once; complete after EnvironmentImpl::synchronize()
    void embedding_nchw_f16_async(uint16_t* out_host);  // the same as fp16 (the chunk's embeddings are converted once, on the work stream)

    int width() const { return size_.orig_w; }
    int height() const { return size_.orig_h; }
    bool encoded() const { return emb_nchw_ != nullptr; }
    EnvironmentImpl& environment() { return env_; }

  private:
    friend class EnvironmentImpl;
    EnvironmentImpl& env_;
    prepost::LongestSide size_;
    std::shared_ptr<StreamBuffer> store_;  // shared by the images of one encoder chunk
    float* emb_nchw_ = nullptr;            // (256, 4096) fp32: the reference's `image_embeddings` layout
    act_t* keys0_ = nullptr;               // (4096, 256) 16-bit: embedding + no_mask_embed, the decoder's layer-0 image stream
    act_t* kvq0_ = nullptr;                // (4096, 384) 16-bit: its layer-0 [K | V | Q] projections
};

}  // namespace dlimg
