// engine.cu -- see engine.hpp.
#include "engine.hpp"
#include "image_pool.hpp"
#include "kernels/mask_select.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <filesystem>

namespace dlimg {

EnvCounters g_unbound_counters;
thread_local EnvCounters* tl_counters = nullptr;

bool pdl_enabled(int family) {
    static unsigned const mask = [] {
        return (unsigned)dev_int("DLIMG_B200_PDL_MASK", 0);  // development switch; measured: no gain inside the CUDA graph (profiles/r01f)
    }();
    return (mask >> family) & 1u;
}

namespace {

constexpr char const* kWeightFile = "mobile_sam_b200.bin";
constexpr size_t kResizedSlot = (size_t)kImageSize * kImageSize * 4;
constexpr size_t kEmbFloats = (size_t)dec::kImgTokens * kEmbedDim;                        // fp32 NCHW embedding
constexpr size_t kKeysElems = (size_t)dec::kImgTokens * 256, kKvqElems = (size_t)dec::kImgTokens * 384;
// per image in a chunk's store: [fp32 embedding | 16-bit keys0 | 16-bit kvq0], every part 256-byte aligned
constexpr size_t kStoreBytesPerImage = kEmbFloats * 4 + (kKeysElems + kKvqElems) * sizeof(act_t);
constexpr size_t kShadowBytesPerImage = kEmbFloats * sizeof(act_t);  // 16-bit copy of the embedding behind a chunk's store

int env_int(char const* name, int def, int lo, int hi) {
    char const* v = std::getenv(name);
    if (!v || !*v) return def;
    int const x = std::atoi(v);
    return std::min(std::max(x, lo), hi);
}

// reference environment.cpp:17-26
std::string verify_model_path(char const* path) {
    namespace fs = std::filesystem;
    std::string const given = path ? path : "models";
    fs::path const p = fs::absolute(given);
    if (!fs::exists(p)) fail("Model path " + given + " does not exist");
    if (!fs::is_directory(p)) fail("Model path " + given + " is not a directory");
    return p.string();
}

void check_view(dlimg_ImageView const& v) {
    if (!v.pixels) fail("Image has no pixel data");
    if (v.width <= 0 || v.height <= 0) fail("Invalid image extent");
    if (!valid_channels(v.channels)) fail("Unsupported channel order [" + std::to_string(v.channels) + "]");
    DLIMG_ASSERT(v.stride >= v.width * bytes_per_pixel(v.channels));  // reference image.cpp:38
}

// prompt assembly, reference segmentation.cpp:134-152: labels 1 / -1 for a point (+ padding point at the transformed
// origin), 2 / 3 for the two box corners; coordinates scaled with round-half-up in float32 (ResizeLongestSide::transform)
void assemble_prompt(dlimg_b200_Prompt const& pr, float scale, float* coords4, float* labels2) {
    if (pr.kind == 0) {
        coords4[0] = float(prepost::scale_coord(pr.x0, scale));
        coords4[1] = float(prepost::scale_coord(pr.y0, scale));
        coords4[2] = float(prepost::scale_coord(0, scale));
        coords4[3] = float(prepost::scale_coord(0, scale));
        labels2[0] = 1.0f;
        labels2[1] = -1.0f;
    } else if (pr.kind == 1) {
        coords4[0] = float(prepost::scale_coord(pr.x0, scale));
        coords4[1] = float(prepost::scale_coord(pr.y0, scale));
        coords4[2] = float(prepost::scale_coord(pr.x1, scale));
        coords4[3] = float(prepost::scale_coord(pr.y1, scale));
        labels2[0] = 2.0f;
        labels2[1] = 3.0f;
    } else {
        fail("Invalid prompt kind " + std::to_string(pr.kind));
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
StreamBuffer::StreamBuffer(size_t bytes, cudaStream_t stream) : stream_(stream) {
    CUDA_CHECK(cudaMallocAsync(&ptr_, bytes, stream));
    CUDA_CHECK(cudaEventCreateWithFlags(&ready_, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&last_read_, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&last_use_, cudaEventDisableTiming));
}
StreamBuffer::~StreamBuffer() {
    // the free is ordered on the allocation stream: downloads (copy-out stream) and decoder passes on a work stream the
    // caller switched to in the meantime may still be reading
    if (read_) cudaStreamWaitEvent(stream_, last_read_, 0);
    if (used_) cudaStreamWaitEvent(stream_, last_use_, 0);
    if (ptr_) cudaFreeAsync(ptr_, stream_);
    if (ready_) cudaEventDestroy(ready_);
    if (last_read_) cudaEventDestroy(last_read_);
    if (last_use_) cudaEventDestroy(last_use_);
    if (f16_ready) cudaEventDestroy(f16_ready);
}
void StreamBuffer::mark_read(cudaStream_t copy_stream) {
    std::lock_guard<std::mutex> lock(mutex_);
    CUDA_CHECK(cudaEventRecord(last_read_, copy_stream));
    read_ = true;
}
void StreamBuffer::mark_used(cudaStream_t work_stream) {
    std::lock_guard<std::mutex> lock(mutex_);
    CUDA_CHECK(cudaEventRecord(last_use_, work_stream));
    used_ = true;
}

// ---------------------------------------------------------------------------------------------
PinnedArena::PinnedArena(size_t bytes) : cap_(bytes) {
    CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&base_), bytes, cudaHostAllocDefault));
}
PinnedArena::~PinnedArena() {
    if (base_) cudaFreeHost(base_);
}
void* PinnedArena::take(size_t bytes, cudaStream_t stream) {
    bytes = (bytes + 63) & ~(size_t)63;
    if (bytes > cap_) fail("pinned arena request too large");
    if (off_ + bytes > cap_) {
        CUDA_CHECK(cudaStreamSynchronize(stream));
        off_ = 0;
    }
    void* p = base_ + off_;
    off_ += bytes;
    return p;
}

// ---------------------------------------------------------------------------------------------
EnvironmentImpl::Scope::Scope(EnvironmentImpl const& env) : prev_counters(tl_counters), prev_profiler(Profiler::bind(&env.profiler_)) {
    tl_counters = &env.counters_;
    env.bind_device();
}
EnvironmentImpl::Scope::~Scope() {
    tl_counters = prev_counters;
    Profiler::bind(prev_profiler);
}

int EnvironmentImpl::device_ordinal() { return env_int("DLIMG_B200_DEVICE", 0, 0, 1023); }

bool EnvironmentImpl::is_supported(dlimg_Backend backend) {
    if (backend != dlimg_gpu) return false;  // there is no CPU path in this engine
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return false;
    }
    int const dev = device_ordinal();
    if (dev >= count) return false;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return prop.major >= 10;  // tcgen05 / TMEM need sm_100
}

EnvironmentImpl::EnvironmentImpl(dlimg_Options const& opts) {
    if (opts.backend != dlimg_gpu)
        fail("Backend::cpu is not available: this build of dlimgedit is the B200-native GPU engine and has no CPU path");
    model_dir_ = verify_model_path(opts.model_directory);
    if (!is_supported(dlimg_gpu)) fail("No CUDA device with compute capability >= 10.0 (Blackwell) is available");
    device_ = device_ordinal();
    CUDA_CHECK(cudaSetDevice(device_));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device_));
    num_sms_ = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&copy_in_, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&copy_out_, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&h2d_done_, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&stream_switch_, cudaEventDisableTiming));
    for (auto& e : input_free_) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&mask_ready_, cudaEventDisableTiming));
    for (auto& e : mask_free_) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // Caps on the images per encoder pass / prompts per decoder pass.  The workspaces behind them are sized by what the
    // calls actually ask for (a single-image, single-prompt caller -- the reference's API -- holds ~120 MB + ~14 MB).
    max_batch_ = env_int("DLIMG_B200_MAX_BATCH", 32, 1, 64);
    max_prompts_ = env_int("DLIMG_B200_MAX_PROMPTS", 64, 1, 256);
    use_graphs_ = env_int("DLIMG_B200_GRAPHS", 1, 0, 1) != 0;
    {   // keep freed stream-ordered allocations cached in the pool instead of returning them to the OS
        cudaMemPool_t pool = nullptr;
        CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device_));
        uint64_t threshold = UINT64_MAX;
        CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    }
    pinned_ = std::make_unique<PinnedArena>((size_t)4 << 20);
    auto const& t = prepost::srgb_tables();
    srgb_decode_.upload(std::vector<float>(t.decode, t.decode + 256));
    srgb_threshold_.upload(std::vector<float>(t.encode_threshold, t.encode_threshold + 256));
    image_pool_attach(device_);  // (last: nothing above may throw after it) façade images are page-locked from now on
}

void EnvironmentImpl::release_graphs() {
    for (auto& g : encode_graphs_) cudaGraphExecDestroy(g.second.exec);
    for (auto& g : decode_graphs_) cudaGraphExecDestroy(g.second.exec);
    encode_graphs_.clear();
    decode_graphs_.clear();
}

EnvironmentImpl::~EnvironmentImpl() {
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    image_pool_detach();
    release_graphs();
    if (own_stream_) cudaStreamDestroy(own_stream_);
    if (copy_in_) cudaStreamDestroy(copy_in_);
    if (copy_out_) cudaStreamDestroy(copy_out_);
    if (h2d_done_) cudaEventDestroy(h2d_done_);
    if (stream_switch_) cudaEventDestroy(stream_switch_);
    for (auto e : input_free_)
        if (e) cudaEventDestroy(e);
    if (mask_ready_) cudaEventDestroy(mask_ready_);
    for (auto e : mask_free_)
        if (e) cudaEventDestroy(e);
}

void EnvironmentImpl::set_stream(cudaStream_t s) {
    std::lock_guard<std::mutex> lock(mutex_);
    Scope scope(*this);
    cudaStream_t const before = stream();
    cudaStream_t const after = s ? s : own_stream_;
    if (before != after) {
        CUDA_CHECK(cudaEventRecord(stream_switch_, before));
        CUDA_CHECK(cudaStreamWaitEvent(after, stream_switch_, 0));
    }
    user_stream_ = s;
}

void EnvironmentImpl::synchronize() {
    bind_device();
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    CUDA_CHECK(cudaStreamSynchronize(copy_in_));
    CUDA_CHECK(cudaStreamSynchronize(copy_out_));
}

void EnvironmentImpl::bind_device() const { CUDA_CHECK(cudaSetDevice(device_)); }

void EnvironmentImpl::profile_enable(bool on) {
    std::lock_guard<std::mutex> lock(mutex_);
    profiler_.enable(on);
}
std::vector<Profiler::Total> EnvironmentImpl::profile_collect() {
    std::lock_guard<std::mutex> lock(mutex_);
    bind_device();
    return profiler_.collect();
}
void EnvironmentImpl::stats(dlimg_b200_Stats& out) const {
    out.kernel_launches = counters_.kernel_launches.load();
    out.h2d_bytes = counters_.h2d_bytes.load();
    out.d2h_bytes = counters_.d2h_bytes.load();
}

SamModel& EnvironmentImpl::model() {
    std::call_once(model_once_, [&] {
        namespace fs = std::filesystem;
        fs::path const dir = fs::path(model_dir_) / "segmentation";
        fs::path const file = dir / kWeightFile;
        if (!fs::exists(file)) {
            std::string msg = "Could not find model file " + file.string();
            if (fs::exists(dir / "mobile_sam_image_encoder.onnx"))
                msg += " (found the reference .onnx files: convert them with tools/convert_checkpoint.py --onnx-dir)";
            fail(msg);
        }
        model_ = std::make_unique<SamModel>(file.string(), num_sms_);
    });
    if (!model_) fail("MobileSAM model failed to load earlier");
    return *model_;
}

// The workspaces grow to the largest pass seen so far.  A CUDA graph holds the workspace addresses it was captured
// with, so growing drops the captured graphs (they are re-captured on next use).
EncoderWorkspace& EnvironmentImpl::encoder_ws(int batch) {
    if (!enc_ws_ || enc_ws_->max_batch < batch) {
        synchronize();
        for (auto& g : encode_graphs_) cudaGraphExecDestroy(g.second.exec);
        encode_graphs_.clear();
        enc_ws_.reset();
        enc_ws_ = std::make_unique<EncoderWorkspace>(batch);
        descs_.allocate((size_t)batch);
    }
    return *enc_ws_;
}
DecoderWorkspace& EnvironmentImpl::decoder_ws(int prompts) {
    if (!dec_ws_ || dec_ws_->max_prompts < prompts) {
        synchronize();
        for (auto& g : decode_graphs_) cudaGraphExecDestroy(g.second.exec);
        decode_graphs_.clear();
        dec_ws_.reset();
        dec_ws_ = std::make_unique<DecoderWorkspace>(prompts);
    }
    return *dec_ws_;
}

DeviceAxisPlan const& EnvironmentImpl::axis_plan(int in_size, int out_size) {
    auto const key = std::make_pair(in_size, out_size);
    auto it = plans_.find(key);
    if (it == plans_.end()) {
        prepost::AxisPlan const p = prepost::make_axis_plan(in_size, out_size);
        DeviceAxisPlan d;
        d.taps = p.taps;
        d.first.upload(p.first);
        d.weights.upload(p.weights);
        d.first_host = p.first;
        it = plans_.emplace(key, std::move(d)).first;
    }
    return it->second;
}

prepost::ResizeDeviceTables EnvironmentImpl::resize_tables(int in_w, int in_h, int out_w, int out_h) {
    DeviceAxisPlan const& hp = axis_plan(in_w, out_w);
    DeviceAxisPlan const& vp = axis_plan(in_h, out_h);
    prepost::ResizeDeviceTables t;
    t.decode = srgb_decode_.get();
    t.encode_threshold = srgb_threshold_.get();
    t.hfirst = hp.first.get(); t.hweights = hp.weights.get(); t.htaps = hp.taps;
    t.vfirst = vp.first.get(); t.vweights = vp.weights.get(); t.vtaps = vp.taps;
    t.hfirst_host = hp.first_host.data();
    t.vfirst_host = vp.first_host.data();
    return t;
}

// Makes one image (already in device memory: the caller's buffer or an uploaded copy) available to the encoder as
// u8 at <= 1024 on the long side.
void EnvironmentImpl::prepare_input(dlimg_ImageView const& view, uint8_t const* dev_pixels, int dev_stride,
                                    prepost::LongestSide const& size, int slot, enc::ImageDesc& desc) {
    cudaStream_t const s = stream();
    int const bpp = bytes_per_pixel(view.channels);
    if (size.needs_resize) {
        prepost::ResizeDeviceTables const t = resize_tables(view.width, view.height, size.w, size.h);
        size_t const need = prepost::resize_scratch_floats(t, view.height, bpp, size.w, size.h);
        if (resize_scratch_.size() < need) {
            CUDA_CHECK(cudaStreamSynchronize(s));
            resize_scratch_.allocate(need);
        }
        uint8_t* dst = resized_px_.get() + (size_t)slot * kResizedSlot;
        prepost::resize_srgb(s, dev_pixels, view.width, view.height, dev_stride, bpp, t, resize_scratch_.get(), dst, size.w, size.h);
        desc.pixels = dst;
        desc.stride = size.w * bpp;
    } else {
        desc.pixels = dev_pixels;
        desc.stride = dev_stride;
    }
    desc._pad = 0;
}

template <typename F> EnvironmentImpl::Graph EnvironmentImpl::capture(cudaStream_t s, F const& body) {
    cudaGraph_t graph = nullptr;
    uint64_t const launches_before = counters_.kernel_launches.load();
    CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    try {
        body();
    } catch (...) {
        cudaStreamEndCapture(s, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
    }
    CUDA_CHECK(cudaStreamEndCapture(s, &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t const err = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_CHECK(err);
    Graph g;
    g.exec = exec;
    g.kernels = counters_.kernel_launches.load() - launches_before;  // kernel nodes recorded by this capture
    counters_.kernel_launches -= g.kernels;                           // they have not run yet
    return g;
}

void EnvironmentImpl::encode_chunk(enc::ImageDesc const* host_descs, int batch, prepost::LongestSide const& size, int channels,
                                   ChunkOut const& out, Tap* tap) {
    cudaStream_t const s = stream();
    SamModel& m = model();
    EncoderWorkspace& ws = encoder_ws(batch);
    size_t const bytes = sizeof(enc::ImageDesc) * (size_t)batch;
    void* staging = pinned_->take(bytes, s);
    std::memcpy(staging, host_descs, bytes);
    CUDA_CHECK(cudaMemcpyAsync(descs_.get(), staging, bytes, cudaMemcpyHostToDevice, s));
    if (tap || !use_graphs_ || profiler_.enabled()) {  // eager launches (debug taps, per-kernel timing)
        m.encode(s, ws, descs_.get(), batch, size.w, size.h, channels, out.emb_nchw, out.keys0, out.kvq0, tap);
        return;
    }
    // The ~70 launches of one encoder pass are captured once per (batch, extent, channel order) into a CUDA
    // graph: every pointer it uses (workspace, weights, descriptor table, tensor maps) is stable, only the
    // descriptor *contents* (uploaded above) and the destination of the embedding change per call -- so the graph ends
    // in front of the final LayerNorm2d, which is launched behind it straight into the caller's store.
    auto const key = std::make_tuple(batch, size.w, size.h, channels);
    auto it = encode_graphs_.find(key);
    if (it == encode_graphs_.end()) {
        Graph const g = capture(s, [&] {
            m.encode(s, ws, descs_.get(), batch, size.w, size.h, channels, nullptr, nullptr, nullptr, nullptr, /*finish=*/false);
        });
        it = encode_graphs_.emplace(key, g).first;
    }
    CUDA_CHECK(cudaGraphLaunch(it->second.exec, s));
    count_launch(it->second.kernels);
    m.neck_finish(s, ws, batch, out.emb_nchw, out.keys0, out.kvq0);
}

void EnvironmentImpl::process_batch(dlimg_ImageView const* views, int count, bool on_device, SegmentationImpl** out) {
    if (count <= 0) return;
    std::lock_guard<std::mutex> lock(mutex_);
    Scope scope(*this);
    for (int i = 0; i < count; ++i) {
        check_view(views[i]);
        if (views[i].width != views[0].width || views[i].height != views[0].height || views[i].channels != views[0].channels)
            fail("process_batch: all images of one call must share width, height and channel order");
    }
    cudaStream_t const s = stream();
    prepost::LongestSide const size = prepost::resize_longest_side(views[0].width, views[0].height, kImageSize);
    int const bpp = bytes_per_pixel(views[0].channels);
    size_t const row_bytes = (size_t)views[0].width * bpp;
    size_t const image_bytes = row_bytes * (size_t)views[0].height;
    int const chunk_cap = std::min(max_batch_, count);
    if (!on_device) {
        size_t const slot = (image_bytes + 255) & ~(size_t)255;
        if (slot > input_slot_bytes_ || (size_t)chunk_cap > input_slots_ || !input_px_[0]) {
            synchronize();
            input_slot_bytes_ = std::max(slot, input_slot_bytes_);
            input_slots_ = std::max((size_t)chunk_cap, input_slots_);
            for (auto& b : input_px_) b.allocate(input_slot_bytes_ * input_slots_);
            input_used_[0] = input_used_[1] = false;
        }
    }
    if (size.needs_resize && (!resized_px_ || resized_slots_ < (size_t)chunk_cap)) {
        synchronize();
        resized_slots_ = (size_t)chunk_cap;
        resized_px_.allocate(kResizedSlot * resized_slots_);
    }

    std::vector<enc::ImageDesc> descs((size_t)chunk_cap);
    for (int start = 0; start < count; start += max_batch_) {
        int const B = std::min(max_batch_, count - start);
        // one allocation per chunk: per image the fp32 NCHW embedding (what get_embedding returns: the reference's
        // `image_embeddings`) and the decoder's 16-bit prompt-independent inputs
        auto store = std::make_shared<StreamBuffer>((kStoreBytesPerImage + kShadowBytesPerImage) * (size_t)B, s);
        store->images = B;
        float* const emb_nchw = reinterpret_cast<float*>(store->bytes());
        act_t* const keys0 = reinterpret_cast<act_t*>(store->bytes() + kEmbFloats * 4 * (size_t)B);
        act_t* const kvq0 = keys0 + kKeysElems * (size_t)B;
        int const set = input_flip_;
        if (!on_device) {
            // upload on the copy stream into slot set `set`, once the encoder that last read it is done; the work
            // stream only waits for this chunk's upload, so uploads run ahead of / alongside earlier encoders.
            // Packed images that follow each other in host memory go up as ONE copy (a batch cut out of one array).
            if (input_used_[set]) CUDA_CHECK(cudaStreamWaitEvent(copy_in_, input_free_[set], 0));
            int i = 0;
            while (i < B) {
                dlimg_ImageView const& v = views[start + i];
                uint8_t* dst = input_px_[set].get() + (size_t)i * input_slot_bytes_;
                bool const packed = (size_t)v.stride == row_bytes;
                int run = 1;
                if (packed && input_slot_bytes_ == image_bytes)
                    while (i + run < B && (size_t)views[start + i + run].stride == row_bytes &&
                           views[start + i + run].pixels == v.pixels + image_bytes * (size_t)run)
                        ++run;
                if (packed) {
                    cudaError_t const merged = cudaMemcpyAsync(dst, v.pixels, image_bytes * (size_t)run, cudaMemcpyHostToDevice, copy_in_);
                    if (merged == cudaErrorInvalidValue && run > 1) {  // separate page-locked allocations that happen to be neighbours
                        cudaGetLastError();
                        for (int r = 0; r < run; ++r)
                            CUDA_CHECK(cudaMemcpyAsync(dst + (size_t)r * image_bytes, v.pixels + (size_t)r * image_bytes, image_bytes,
                                                       cudaMemcpyHostToDevice, copy_in_));
                    } else {
                        CUDA_CHECK(merged);
                    }
                } else {
                    CUDA_CHECK(cudaMemcpy2DAsync(dst, row_bytes, v.pixels, (size_t)v.stride, row_bytes, (size_t)v.height,
                                                 cudaMemcpyHostToDevice, copy_in_));
                }
                counters_.h2d_bytes += image_bytes * (size_t)run;
                i += run;
            }
            CUDA_CHECK(cudaEventRecord(h2d_done_, copy_in_));
            CUDA_CHECK(cudaStreamWaitEvent(s, h2d_done_, 0));
        }
        for (int i = 0; i < B; ++i) {
            dlimg_ImageView const& v = views[start + i];
            if (on_device) prepare_input(v, v.pixels, v.stride, size, i, descs[(size_t)i]);
            else prepare_input(v, input_px_[set].get() + (size_t)i * input_slot_bytes_, (int)row_bytes, size, i, descs[(size_t)i]);
        }
        encode_chunk(descs.data(), B, size, views[0].channels, ChunkOut{emb_nchw, keys0, kvq0}, nullptr);
        CUDA_CHECK(cudaEventRecord(store->ready(), s));
        if (!on_device) {
            CUDA_CHECK(cudaEventRecord(input_free_[set], s));
            input_used_[set] = true;
            input_flip_ ^= 1;
        }
        for (int i = 0; i < B; ++i) {
            SegmentationImpl* seg = out[start + i];
            seg->size_ = size;
            seg->store_ = store;
            seg->emb_nchw_ = emb_nchw + (size_t)i * kEmbFloats;
            seg->keys0_ = keys0 + (size_t)i * kKeysElems;
            seg->kvq0_ = kvq0 + (size_t)i * kKvqElems;
        }
    }
    // host pixels are only borrowed for the duration of the call: wait for the uploads (not for the encoder)
    if (!on_device) CUDA_CHECK(cudaStreamSynchronize(copy_in_));
}

// One decoder pass over P prompts, which may belong to different images: uploads the parameter block (coordinates,
// labels, per-prompt image tables) in one copy and launches the pass -- a CUDA graph per P, since every address in it is
// fixed (workspace, weights, parameter block) and only the block's contents change.
void EnvironmentImpl::decode_chunk(SegmentationImpl* const* segs, dlimg_b200_Prompt const* prompts, int P, bool eager, int mask_mode) {
    cudaStream_t const s = stream();
    SamModel& m = model();
    DecoderWorkspace& ws = decoder_ws(P);
    size_t const bytes = DecoderParams::bytes(P);
    uint8_t* staging = static_cast<uint8_t*>(pinned_->take(bytes, s));
    DecoderParams const h = DecoderWorkspace::layout(staging, P);
    StreamBuffer* last_store = nullptr;
    for (int p = 0; p < P; ++p) {
        SegmentationImpl* seg = segs[p];
        if (!seg || !seg->encoded()) fail("compute_mask: segmentation handle holds no processed image");
        if (&seg->env_ != this) fail("compute_mask: segmentation handle belongs to another environment");
        assemble_prompt(prompts[p], seg->size_.scale, h.coords + 4 * p, h.labels + 2 * p);
        h.keys0[p] = seg->keys0_;
        h.kvq0[p] = seg->kvq0_;
        if (seg->store_.get() != last_store) {  // the encoder may have run on another stream (set_stream)
            last_store = seg->store_.get();
            CUDA_CHECK(cudaStreamWaitEvent(s, last_store->ready(), 0));
        }
    }
    CUDA_CHECK(cudaMemcpyAsync(ws.param_block.get(), staging, bytes, cudaMemcpyHostToDevice, s));
    counters_.h2d_bytes += bytes;
    if (eager || !use_graphs_ || profiler_.enabled()) {
        m.decode(s, ws, P, mask_mode);
    } else {
        int const key = P * 4 + mask_mode;
        auto it = decode_graphs_.find(key);
        if (it == decode_graphs_.end()) it = decode_graphs_.emplace(key, capture(s, [&] { m.decode(s, ws, P, mask_mode); })).first;
        CUDA_CHECK(cudaGraphLaunch(it->second.exec, s));
        count_launch(it->second.kernels);
    }
    last_store = nullptr;
    for (int p = 0; p < P; ++p)
        if (segs[p]->store_.get() != last_store) {
            last_store = segs[p]->store_.get();
            last_store->mark_used(s);
        }
}

void EnvironmentImpl::compute_masks_batch(SegmentationImpl* const* segs, dlimg_b200_Prompt const* prompts, int count, bool multi,
                                          uint8_t* const* planes_out, float* ious_out, int placement) {
    if (count <= 0) return;
    bool const on_device = placement == 1, async_host = placement == 2;
    std::lock_guard<std::mutex> lock(mutex_);
    Scope scope(*this);
    cudaStream_t const s = stream();
    int const n = multi ? 3 : 1;
    // host-destined masks go in groups of at most kHostGroup prompts so that downloads overlap decoding (splitting a call
    // into smaller decoder groups costs more than it hides -- 18.3k -> 9.5k masks/s at 16 -- so the default keeps whole groups)
    int const kHostGroup = env_int("DLIMG_B200_HOST_GROUP", 256, 1, 256);
    int const group_cap = on_device ? max_prompts_ : std::min(max_prompts_, kHostGroup);
    for (int i = 0; i < count; i += group_cap) {
        int const P = std::min(group_cap, count - i);
        decode_chunk(segs + i, prompts + i, P, false, multi ? MASKS_MULTI : MASKS_BEST);
        DecoderWorkspace& ws = *dec_ws_;
        int const planes = P * n;
        float* iou_dst = (on_device && ious_out) ? ious_out + (size_t)i * n : ws.iou_sel.get();
        dec::select_masks(s, ws.iou.get(), P, multi ? 1 : 0, ws.plane_index.get(), iou_dst);

        // destination of every plane: the caller's device buffers, or a contiguous device staging area for host callers
        size_t total_bytes = 0;
        for (int p = 0; p < P; ++p) total_bytes += (size_t)segs[i + p]->width() * segs[i + p]->height() * n;
        uint8_t* staging_dev = nullptr;
        int flip = 0;
        if (!on_device) {
            flip = host_group_ & 1;
            ++host_group_;
            if (mask_out_[flip].size() < total_bytes) {
                synchronize();
                mask_out_[flip].allocate(total_bytes);
                mask_used_[flip] = false;
            }
            if (mask_used_[flip]) CUDA_CHECK(cudaStreamWaitEvent(s, mask_free_[flip], 0));  // its previous download is done
            staging_dev = mask_out_[flip].get();
        }
        if (plane_ptrs_.size() < (size_t)planes) {
            CUDA_CHECK(cudaStreamSynchronize(s));
            plane_ptrs_.allocate((size_t)std::max(planes, max_prompts_ * 3));
        }
        uint8_t** hp = static_cast<uint8_t**>(pinned_->take(sizeof(uint8_t*) * (size_t)planes, s));
        {
            size_t off = 0;
            for (int p = 0; p < P; ++p) {
                size_t const plane_bytes = (size_t)segs[i + p]->width() * segs[i + p]->height();
                for (int k = 0; k < n; ++k) {
                    hp[p * n + k] = on_device ? planes_out[(size_t)(i + p) * n + k] : staging_dev + off;
                    off += plane_bytes;
                }
            }
        }
        CUDA_CHECK(cudaMemcpyAsync(plane_ptrs_.get(), hp, sizeof(uint8_t*) * (size_t)planes, cudaMemcpyHostToDevice, s));
        // one post-processing launch per run of prompts whose images share their geometry
        for (int p0 = 0; p0 < P;) {
            prepost::LongestSide const& g = segs[i + p0]->size_;
            int p1 = p0 + 1;
            while (p1 < P && segs[i + p1]->size_.orig_w == g.orig_w && segs[i + p1]->size_.orig_h == g.orig_h) ++p1;
            prepost::mask_postprocess(s, ws.low.get(), 65536, ws.plane_index.get() + p0 * n, (p1 - p0) * n, g.w, g.h, g.orig_w, g.orig_h,
                                      plane_ptrs_.get() + p0 * n);
            p0 = p1;
        }
        if (!on_device) {
            if (ious_out) {  // tiny, and ws.iou_sel is rewritten by the next group: keep it on the work stream
                CUDA_CHECK(cudaMemcpyAsync(ious_out + (size_t)i * n, ws.iou_sel.get(), sizeof(float) * (size_t)planes,
                                           cudaMemcpyDeviceToHost, s));
                counters_.d2h_bytes += sizeof(float) * (size_t)planes;
            }
            CUDA_CHECK(cudaEventRecord(mask_ready_, s));
            CUDA_CHECK(cudaStreamWaitEvent(copy_out_, mask_ready_, 0));
            // planes that follow each other in the caller's memory leave as one copy
            int k = 0;
            while (k < planes) {
                size_t const first_off = (size_t)(hp[k] - staging_dev);
                uint8_t* const host_dst = planes_out[(size_t)i * n + k];
                size_t run_bytes = 0;
                int k1 = k;
                while (k1 < planes && planes_out[(size_t)i * n + k1] == host_dst + run_bytes) {
                    SegmentationImpl const* sg = segs[i + k1 / n];
                    run_bytes += (size_t)sg->width() * sg->height();
                    ++k1;
                }
                cudaError_t const merged = cudaMemcpyAsync(host_dst, staging_dev + first_off, run_bytes, cudaMemcpyDeviceToHost, copy_out_);
                if (merged == cudaErrorInvalidValue && k1 - k > 1) {
                    // neighbouring buffers that are SEPARATE page-locked allocations: one copy may not span them
                    cudaGetLastError();
                    size_t off = 0;
                    for (int kk = k; kk < k1; ++kk) {
                        SegmentationImpl const* sg = segs[i + kk / n];
                        size_t const bytes = (size_t)sg->width() * sg->height();
                        CUDA_CHECK(cudaMemcpyAsync(host_dst + off, staging_dev + first_off + off, bytes, cudaMemcpyDeviceToHost, copy_out_));
                        off += bytes;
                    }
                } else {
                    CUDA_CHECK(merged);
                }
                k = k1;
            }
            counters_.d2h_bytes += total_bytes;
            CUDA_CHECK(cudaEventRecord(mask_free_[flip], copy_out_));
            mask_used_[flip] = true;
        }
    }
    if (!on_device && !async_host) {  // the caller's host buffers are complete when the call returns
        CUDA_CHECK(cudaStreamSynchronize(s));
        CUDA_CHECK(cudaStreamSynchronize(copy_out_));
    }
}

void EnvironmentImpl::low_res_logits(SegmentationImpl& seg, dlimg_b200_Prompt const& prompt, float* logits_host, float* iou_host) {
    // decode one prompt, then read back the unselected (4, 256, 256) logits and (4) IoU predictions
    std::lock_guard<std::mutex> lock(mutex_);
    Scope scope(*this);
    cudaStream_t const s = stream();
    SegmentationImpl* segs[1] = {&seg};
    decode_chunk(segs, &prompt, 1, true, MASKS_ALL);
    DecoderWorkspace& ws = *dec_ws_;
    CUDA_CHECK(cudaMemcpyAsync(logits_host, ws.low.get(), sizeof(float) * 4 * 65536, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(iou_host, ws.iou.get(), sizeof(float) * 4, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
}

// ---------------------------------------------------------------------------------------------
void EnvironmentImpl::resize_longest_side(dlimg_ImageView const& v, int max_side, uint8_t* dev_out, int* out_extent) {
    std::lock_guard<std::mutex> lock(mutex_);
    Scope scope(*this);
    check_view(v);
    cudaStream_t const s = stream();
    prepost::LongestSide const size = prepost::resize_longest_side(v.width, v.height, max_side);
    int const bpp = bytes_per_pixel(v.channels);
    if (out_extent) { out_extent[0] = size.w; out_extent[1] = size.h; }
    if (!size.needs_resize) {  // the reference returns the original view (segmentation.cpp:63-70); emit a packed copy
        CUDA_CHECK(cudaMemcpy2DAsync(dev_out, (size_t)v.width * bpp, v.pixels, (size_t)v.stride, (size_t)v.width * bpp,
                                     (size_t)v.height, cudaMemcpyDeviceToDevice, s));
        return;
    }
    prepost::ResizeDeviceTables const t = resize_tables(v.width, v.height, size.w, size.h);
    size_t const need = prepost::resize_scratch_floats(t, v.height, bpp, size.w, size.h);
    if (resize_scratch_.size() < need) {
        CUDA_CHECK(cudaStreamSynchronize(s));
        resize_scratch_.allocate(need);
    }
    prepost::resize_srgb(s, v.pixels, v.width, v.height, v.stride, bpp, t, resize_scratch_.get(), dev_out, size.w, size.h);
}

void EnvironmentImpl::image_tensor(dlimg_ImageView const& v, float* dev_out) {
    std::lock_guard<std::mutex> lock(mutex_);
    Scope scope(*this);
    check_view(v);
    prepost::image_tensor(stream(), v.pixels, v.width, v.height, v.stride, v.channels, dev_out);
}

void EnvironmentImpl::mask_postprocess(float const* dev_low_res, int count, int w, int h, uint8_t* dev_out) {
    std::lock_guard<std::mutex> lock(mutex_);
    Scope scope(*this);
    prepost::LongestSide const size = prepost::resize_longest_side(w, h, kImageSize);
    prepost::mask_postprocess_contiguous(stream(), dev_low_res, 65536, nullptr, count, size.w, size.h, w, h, dev_out);
}

void EnvironmentImpl::threshold_mask(float const* dev_logits, int th, int tw, int w, int h, uint8_t* dev_out) {
    std::lock_guard<std::mutex> lock(mutex_);
    Scope scope(*this);
    prepost::threshold_mask(stream(), dev_logits, th, tw, w, h, dev_out);
}

size_t EnvironmentImpl::encode_tap(dlimg_ImageView const* views, int count, char const* tap_name, float* dev_out, size_t capacity) {
    std::lock_guard<std::mutex> lock(mutex_);
    Scope scope(*this);
    DLIMG_ASSERT(count >= 1 && count <= max_batch_);
    for (int i = 0; i < count; ++i) check_view(views[i]);
    prepost::LongestSide const size = prepost::resize_longest_side(views[0].width, views[0].height, kImageSize);
    if (size.needs_resize && (!resized_px_ || resized_slots_ < (size_t)count)) {
        synchronize();
        resized_slots_ = (size_t)count;
        resized_px_.allocate(kResizedSlot * resized_slots_);
    }
    std::vector<enc::ImageDesc> descs((size_t)count);
    for (int i = 0; i < count; ++i) prepare_input(views[i], views[i].pixels, views[i].stride, size, i, descs[(size_t)i]);
    DeviceBuffer<uint8_t> store(kStoreBytesPerImage * (size_t)count);
    float* const emb_nchw = reinterpret_cast<float*>(store.get());
    act_t* const keys0 = reinterpret_cast<act_t*>(store.get() + kEmbFloats * 4 * (size_t)count);
    Tap tap;
    tap.name = tap_name;
    tap.out = dev_out;
    tap.capacity = capacity;
    encode_chunk(descs.data(), count, size, views[0].channels, ChunkOut{emb_nchw, keys0, keys0 + kKeysElems * (size_t)count}, &tap);
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    return tap.written;
}

// ---------------------------------------------------------------------------------------------
void SegmentationImpl::process(dlimg_ImageView const& view) {
    SegmentationImpl* self = this;
    env_.process_batch(&view, 1, false, &self);
}

void SegmentationImpl::compute_mask(int const* point, int const* region, uint8_t** out_masks, float* out_accuracy) {
    DLIMG_ASSERT(point || region);  // reference segmentation.cpp:134
    dlimg_b200_Prompt pr{};
    if (point) {
        pr.kind = 0;
        pr.x0 = point[0];
        pr.y0 = point[1];
    } else {
        pr.kind = 1;
        pr.x0 = region[0];
        pr.y0 = region[1];
        pr.x1 = region[2];
        pr.y1 = region[3];
    }
    SegmentationImpl* self = this;
    bool const is_single = out_masks[1] == nullptr;  // reference segmentation.cpp:154
    if (is_single) {
        DLIMG_ASSERT(out_masks[0] != nullptr);
        env_.compute_masks_batch(&self, &pr, 1, false, out_masks, nullptr, 0);  // accuracy is not written (:162-165)
    } else {
        for (int i = 0; i < 3; ++i) DLIMG_ASSERT(out_masks[i] != nullptr);
        env_.compute_masks_batch(&self, &pr, 1, true, out_masks, out_accuracy, 0);
    }
}

// The encoder leaves the fp32 NCHW embedding in the chunk's store; the device->host transfer runs on the copy-out
// stream as soon as that chunk is ready, so it overlaps whatever the work stream does next.
void SegmentationImpl::embedding_nchw_async(float* out_host) {
    std::lock_guard<std::mutex> lock(env_.mutex());
    EnvironmentImpl::Scope scope(env_);
    if (!encoded()) fail("segmentation handle holds no processed image");
    CUDA_CHECK(cudaStreamWaitEvent(env_.copy_out_, store_->ready(), 0));
    CUDA_CHECK(cudaMemcpyAsync(out_host, emb_nchw_, kEmbFloats * sizeof(float), cudaMemcpyDeviceToHost, env_.copy_out_));
    store_->mark_read(env_.copy_out_);
    env_.counters_.d2h_bytes += kEmbFloats * sizeof(float);
}

void SegmentationImpl::embedding_nchw_f16_async(uint16_t* out_host) {
#if defined(DLIMG_B200_ACT_BF16)
    (void)out_host;
    fail("get_embedding_f16_async: this build stores bf16");
#else
    std::lock_guard<std::mutex> lock(env_.mutex());
    EnvironmentImpl::Scope scope(env_);
    if (!encoded()) fail("segmentation handle holds no processed image");
    // The conversion runs on the WORK stream, in order behind the encoder (on the copy-out stream it had to squeeze onto SMs
    // that the next pass's persistent kernels occupy, and the downloads behind it stalled unpredictably), ONCE for all images
    // of the chunk into the 16-bit region behind the chunk's store: one launch per chunk instead of one per image, and no
    // per-handle allocation (a fresh 2 MiB block per new handle made the memory pool grow in the middle of a pipelined run).
    // The copy follows on the copy-out stream.
    cudaStream_t const ws = env_.stream(), cs = env_.copy_out_;
    StreamBuffer& st = *store_;
    float const* const emb_base = reinterpret_cast<float const*>(st.bytes());
    act_t* const shadow_base = reinterpret_cast<act_t*>(st.bytes() + kStoreBytesPerImage * (size_t)st.images);
    if (!st.f16_done) {
        CUDA_CHECK(cudaStreamWaitEvent(ws, st.ready(), 0));
        if (!st.f16_ready) CUDA_CHECK(cudaEventCreateWithFlags(&st.f16_ready, cudaEventDisableTiming));
        dec::f32_to_act(ws, emb_base, (int64_t)kEmbFloats * st.images, shadow_base);
        CUDA_CHECK(cudaEventRecord(st.f16_ready, ws));
        st.mark_used(ws);
        st.f16_done = true;
    }
    size_t const index = (size_t)(emb_nchw_ - emb_base) / kEmbFloats;
    CUDA_CHECK(cudaStreamWaitEvent(cs, st.f16_ready, 0));
    CUDA_CHECK(cudaMemcpyAsync(out_host, shadow_base + index * kEmbFloats, kEmbFloats * sizeof(act_t), cudaMemcpyDeviceToHost, cs));
    st.mark_read(cs);
    env_.counters_.d2h_bytes += kEmbFloats * sizeof(act_t);
#endif
}

void SegmentationImpl::embedding_nchw(float* out_host) {
    embedding_nchw_async(out_host);
    CUDA_CHECK(cudaStreamSynchronize(env_.copy_out_));
}

}  // namespace dlimg
