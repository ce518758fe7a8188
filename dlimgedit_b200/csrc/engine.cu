// engine.cu -- see engine.hpp.
#include "engine.hpp"

#include "profiler.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <filesystem>

namespace dlimg {

std::atomic<uint64_t> g_kernel_launches{0};
bool pdl_enabled(int family) {
    static unsigned const mask = [] {
        char const* e = std::getenv("DLIMG_B200_PDL_MASK");
        return e ? (unsigned)std::strtoul(e, nullptr, 0) : 0u;  // measured: no gain inside the CUDA graph (profiles/r01f)
    }();
    return (mask >> family) & 1u;
}
std::atomic<uint64_t> g_h2d_bytes{0};
std::atomic<uint64_t> g_d2h_bytes{0};

namespace {

constexpr char const* kWeightFile = "mobile_sam_b200.bin";
constexpr size_t kResizedSlot = (size_t)kImageSize * kImageSize * 4;

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

}  // namespace

// ---------------------------------------------------------------------------------------------
StreamBuffer::StreamBuffer(size_t bytes, cudaStream_t stream) : stream_(stream) {
    CUDA_CHECK(cudaMallocAsync(&ptr_, bytes, stream));
    CUDA_CHECK(cudaEventCreateWithFlags(&ready_, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&last_read_, cudaEventDisableTiming));
}
StreamBuffer::~StreamBuffer() {
    if (read_) cudaStreamWaitEvent(stream_, last_read_, 0);  // a download on the copy-out stream may still be in flight
    if (ptr_) cudaFreeAsync(ptr_, stream_);
    if (ready_) cudaEventDestroy(ready_);
    if (last_read_) cudaEventDestroy(last_read_);
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
    CUDA_CHECK(cudaEventCreateWithFlags(&d2h_ready_, cudaEventDisableTiming));
    for (auto& e : input_free_) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&mask_ready_, cudaEventDisableTiming));
    for (auto& e : mask_free_) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    max_batch_ = env_int("DLIMG_B200_MAX_BATCH", 8, 1, 64);
    max_prompts_ = env_int("DLIMG_B200_MAX_PROMPTS", 32, 1, 256);
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
}

EnvironmentImpl::~EnvironmentImpl() {
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    for (auto& g : encode_graphs_) cudaGraphExecDestroy(g.second.exec);
    if (own_stream_) cudaStreamDestroy(own_stream_);
    if (copy_in_) cudaStreamDestroy(copy_in_);
    if (copy_out_) cudaStreamDestroy(copy_out_);
    if (h2d_done_) cudaEventDestroy(h2d_done_);
    if (d2h_ready_) cudaEventDestroy(d2h_ready_);
    for (auto e : input_free_)
        if (e) cudaEventDestroy(e);
    if (mask_ready_) cudaEventDestroy(mask_ready_);
    for (auto e : mask_free_)
        if (e) cudaEventDestroy(e);
}

void EnvironmentImpl::synchronize() {
    bind_device();
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    CUDA_CHECK(cudaStreamSynchronize(copy_in_));
    CUDA_CHECK(cudaStreamSynchronize(copy_out_));
}

void EnvironmentImpl::bind_device() const { CUDA_CHECK(cudaSetDevice(device_)); }

SamModel& EnvironmentImpl::model() {
    std::call_once(model_once_, [&] {
        namespace fs = std::filesystem;
        fs::path const dir = fs::path(model_dir_) / "segmentation";
        fs::path const file = dir / kWeightFile;
        if (!fs::exists(file)) {
            std::string msg = "Could not find model file " + file.string();
            if (fs::exists(dir / "mobile_sam_image_encoder.onnx"))
                msg += " (found the reference .onnx files: convert the MobileSAM checkpoint with tools/convert_checkpoint.py)";
            fail(msg);
        }
        model_ = std::make_unique<SamModel>(file.string(), num_sms_);
    });
    if (!model_) fail("MobileSAM model failed to load earlier");
    return *model_;
}

EncoderWorkspace& EnvironmentImpl::encoder_ws() {
    if (!enc_ws_) enc_ws_ = std::make_unique<EncoderWorkspace>(max_batch_);
    return *enc_ws_;
}
DecoderWorkspace& EnvironmentImpl::decoder_ws() {
    if (!dec_ws_) dec_ws_ = std::make_unique<DecoderWorkspace>(max_prompts_);
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
    return t;
}

// Makes one image (already in device memory: the caller's buffer or an uploaded copy) available to the encoder as
// u8 at <= 1024 on the long side.
uint8_t* EnvironmentImpl::prepare_input(dlimg_ImageView const& view, uint8_t const* dev_pixels, int dev_stride,
                                        prepost::LongestSide const& size, int slot, enc::ImageDesc& desc) {
    cudaStream_t const s = stream();
    int const bpp = bytes_per_pixel(view.channels);
    uint8_t const* src = dev_pixels;
    int const src_stride = dev_stride;
    if (size.needs_resize) {
        size_t const need = (size_t)view.height * size.w * bpp;
        if (resize_scratch_.size() < need) {
            CUDA_CHECK(cudaStreamSynchronize(s));
            resize_scratch_.allocate(need);
        }
        uint8_t* dst = resized_px_.get() + (size_t)slot * kResizedSlot;
        prepost::resize_srgb(s, src, view.width, view.height, src_stride, bpp,
                             resize_tables(view.width, view.height, size.w, size.h), resize_scratch_.get(), dst, size.w, size.h);
        desc.pixels = dst;
        desc.stride = size.w * bpp;
    } else {
        desc.pixels = src;
        desc.stride = src_stride;
    }
    desc._pad = 0;
    return const_cast<uint8_t*>(desc.pixels);
}

void EnvironmentImpl::encode_chunk(enc::ImageDesc const* host_descs, int batch, prepost::LongestSide const& size, int channels,
                                   float* emb_out, Tap* tap, float* emb_nchw_out) {
    cudaStream_t const s = stream();
    if (!descs_) descs_.allocate((size_t)max_batch_);
    size_t const bytes = sizeof(enc::ImageDesc) * (size_t)batch;
    void* staging = pinned_->take(bytes, s);
    std::memcpy(staging, host_descs, bytes);
    CUDA_CHECK(cudaMemcpyAsync(descs_.get(), staging, bytes, cudaMemcpyHostToDevice, s));
    SamModel& m = model();
    EncoderWorkspace& ws = encoder_ws();
    if (tap || !use_graphs_ || Profiler::get().enabled()) {  // eager launches (debug taps, per-kernel timing)
        m.encode(s, ws, descs_.get(), batch, size.w, size.h, channels, emb_out, tap, emb_nchw_out);
        return;
    }
    // The ~70 launches of one encoder pass are captured once per (batch, extent, channel order) into a CUDA
    // graph: every pointer it uses (workspace, weights, descriptor table, tensor maps) is stable, only the
    // descriptor *contents* (uploaded above) and the destination of the embedding change per call -- so the graph ends
    // in front of the final LayerNorm2d, which is launched behind it straight into the caller's store (both layouts).
    auto const key = std::make_tuple(batch, size.w, size.h, channels);
    auto it = encode_graphs_.find(key);
    if (it == encode_graphs_.end()) {
        cudaGraph_t graph = nullptr;
        uint64_t const launches_before = g_kernel_launches.load();
        CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        try {
            m.encode(s, ws, descs_.get(), batch, size.w, size.h, channels, nullptr, nullptr, nullptr, /*finish=*/false);
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
        EncodeGraph eg;
        eg.exec = exec;
        eg.kernels = g_kernel_launches.load() - launches_before;  // kernel nodes recorded by this capture
        g_kernel_launches -= eg.kernels;                          // they have not run yet
        it = encode_graphs_.emplace(key, eg).first;
    }
    CUDA_CHECK(cudaGraphLaunch(it->second.exec, s));
    count_launch(it->second.kernels);
    m.neck_finish(s, ws, batch, emb_out, emb_nchw_out);
}

void EnvironmentImpl::process_batch(dlimg_ImageView const* views, int count, bool on_device, SegmentationImpl** out) {
    if (count <= 0) return;
    std::lock_guard<std::mutex> lock(mutex_);
    bind_device();
    for (int i = 0; i < count; ++i) {
        check_view(views[i]);
        if (views[i].width != views[0].width || views[i].height != views[0].height || views[i].channels != views[0].channels)
            fail("process_batch: all images of one call must share width, height and channel order");
    }
    cudaStream_t const s = stream();
    prepost::LongestSide const size = prepost::resize_longest_side(views[0].width, views[0].height, kImageSize);
    int const bpp = bytes_per_pixel(views[0].channels);
    size_t const row_bytes = (size_t)views[0].width * bpp;
    if (!on_device) {
        size_t const slot = ((size_t)views[0].width * views[0].height * bpp + 255) & ~(size_t)255;
        if (slot > input_slot_bytes_ || !input_px_[0]) {
            synchronize();
            input_slot_bytes_ = slot;
            for (auto& b : input_px_) b.allocate(slot * (size_t)max_batch_);
            input_used_[0] = input_used_[1] = false;
        }
    }
    if (size.needs_resize && !resized_px_) resized_px_.allocate(kResizedSlot * (size_t)max_batch_);

    std::vector<enc::ImageDesc> descs((size_t)max_batch_);
    for (int start = 0; start < count; start += max_batch_) {
        int const B = std::min(max_batch_, count - start);
        // one allocation per chunk: B token-major embeddings (what the decoder reads) + the same in NCHW (what
        // get_embedding returns: the reference's `image_embeddings` layout), so a read is a plain copy
        size_t const emb_floats = (size_t)dec::kImgTokens * kEmbedDim;
        auto store = std::make_shared<StreamBuffer>(sizeof(float) * 2 * (size_t)B * emb_floats, s);
        int const set = input_flip_;
        if (!on_device) {
            // upload on the copy stream into slot set `set`, once the encoder that last read it is done; the work
            // stream only waits for this chunk's upload, so uploads run ahead of / alongside earlier encoders
            if (input_used_[set]) CUDA_CHECK(cudaStreamWaitEvent(copy_in_, input_free_[set], 0));
            for (int i = 0; i < B; ++i) {
                dlimg_ImageView const& v = views[start + i];
                uint8_t* dst = input_px_[set].get() + (size_t)i * input_slot_bytes_;
                CUDA_CHECK(cudaMemcpy2DAsync(dst, row_bytes, v.pixels, (size_t)v.stride, row_bytes, (size_t)v.height,
                                             cudaMemcpyHostToDevice, copy_in_));
                g_h2d_bytes += row_bytes * (size_t)v.height;
            }
            CUDA_CHECK(cudaEventRecord(h2d_done_, copy_in_));
            CUDA_CHECK(cudaStreamWaitEvent(s, h2d_done_, 0));
        }
        for (int i = 0; i < B; ++i) {
            dlimg_ImageView const& v = views[start + i];
            if (on_device) prepare_input(v, v.pixels, v.stride, size, i, descs[(size_t)i]);
            else prepare_input(v, input_px_[set].get() + (size_t)i * input_slot_bytes_, (int)row_bytes, size, i, descs[(size_t)i]);
        }
        encode_chunk(descs.data(), B, size, views[0].channels, store->floats(), nullptr, store->floats() + (size_t)B * emb_floats);
        CUDA_CHECK(cudaEventRecord(store->ready(), s));
        if (!on_device) {
            CUDA_CHECK(cudaEventRecord(input_free_[set], s));
            input_used_[set] = true;
            input_flip_ ^= 1;
        }
        for (int i = 0; i < B; ++i) {
            SegmentationImpl* seg = out[start + i];
            seg->size_ = size;
            seg->emb_store_ = store;
            seg->emb_ = store->floats() + (size_t)i * emb_floats;
            seg->emb_nchw_ = store->floats() + ((size_t)B + i) * emb_floats;
            seg->cache_.ready = false;
        }
    }
    // host pixels are only borrowed for the duration of the call: wait for the uploads (not for the encoder)
    if (!on_device) CUDA_CHECK(cudaStreamSynchronize(copy_in_));
}

void EnvironmentImpl::compute_masks_batch(SegmentationImpl* const* segs, dlimg_b200_Prompt const* prompts, int count, bool multi,
                                          uint8_t* const* planes_out, float* ious_out, int placement) {
    if (count <= 0) return;
    bool const on_device = placement == 1, async_host = placement == 2;
    std::lock_guard<std::mutex> lock(mutex_);
    bind_device();
    cudaStream_t const s = stream();
    SamModel& m = model();
    DecoderWorkspace& ws = decoder_ws();
    int const n = multi ? 3 : 1;
    int const kHostGroup = env_int("DLIMG_B200_HOST_GROUP", 256, 1, 256);  // splitting a call into smaller decoder groups to overlap downloads costs more than it hides (18.3k -> 9.5k masks/s at 16), so the default keeps whole groups
    int i = 0;
    while (i < count) {
        SegmentationImpl* seg = segs[i];
        if (!seg || !seg->encoded()) fail("compute_mask: segmentation handle holds no processed image");
        int j = i;
        // host-destined masks go in groups of at most kHostGroup prompts so that downloads overlap decoding
        int const group_cap = on_device ? max_prompts_ : std::min(max_prompts_, kHostGroup);
        while (j < count && segs[j] == seg && j - i < group_cap) ++j;
        int const P = j - i;
        if (!seg->cache_.ready) m.prepare_embedding(s, seg->emb_, seg->cache_);

        // prompt assembly, reference segmentation.cpp:134-152: labels 1 / -1 for a point (+ padding point at the
        // transformed origin), 2 / 3 for the two box corners; coordinates scaled with round-half-up in float32
        float* hc = static_cast<float*>(pinned_->take(sizeof(float) * 6 * (size_t)P, s));
        float* hl = hc + 4 * P;
        float const scale = seg->size_.scale;
        for (int p = 0; p < P; ++p) {
            dlimg_b200_Prompt const& pr = prompts[i + p];
            if (pr.kind == 0) {
                hc[4 * p + 0] = float(prepost::scale_coord(pr.x0, scale));
                hc[4 * p + 1] = float(prepost::scale_coord(pr.y0, scale));
                hc[4 * p + 2] = float(prepost::scale_coord(0, scale));
                hc[4 * p + 3] = float(prepost::scale_coord(0, scale));
                hl[2 * p + 0] = 1.0f;
                hl[2 * p + 1] = -1.0f;
            } else if (pr.kind == 1) {
                hc[4 * p + 0] = float(prepost::scale_coord(pr.x0, scale));
                hc[4 * p + 1] = float(prepost::scale_coord(pr.y0, scale));
                hc[4 * p + 2] = float(prepost::scale_coord(pr.x1, scale));
                hc[4 * p + 3] = float(prepost::scale_coord(pr.y1, scale));
                hl[2 * p + 0] = 2.0f;
                hl[2 * p + 1] = 3.0f;
            } else {
                fail("Invalid prompt kind " + std::to_string(pr.kind));
            }
        }
        CUDA_CHECK(cudaMemcpyAsync(ws.coords.get(), hc, sizeof(float) * 4 * (size_t)P, cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaMemcpyAsync(ws.labels.get(), hl, sizeof(float) * 2 * (size_t)P, cudaMemcpyHostToDevice, s));
        g_h2d_bytes += sizeof(float) * 6 * (size_t)P;

        m.decode(s, ws, seg->cache_, P);

        int const W = seg->width(), H = seg->height();
        size_t const plane_bytes = (size_t)W * H;
        int const planes = P * n;
        float* iou_dst = (on_device && ious_out) ? ious_out + (size_t)i * n : ws.iou_sel.get();
        dec::select_masks(s, ws.iou.get(), P, multi ? 1 : 0, ws.plane_index.get(), iou_dst);
        if (on_device) {
            if (plane_ptrs_.size() < (size_t)planes) {
                CUDA_CHECK(cudaStreamSynchronize(s));
                plane_ptrs_.allocate((size_t)std::max(planes, max_prompts_ * 3));
            }
            uint8_t** hp = static_cast<uint8_t**>(pinned_->take(sizeof(uint8_t*) * (size_t)planes, s));
            for (int k = 0; k < planes; ++k) hp[k] = planes_out[(size_t)i * n + k];
            CUDA_CHECK(cudaMemcpyAsync(plane_ptrs_.get(), hp, sizeof(uint8_t*) * (size_t)planes, cudaMemcpyHostToDevice, s));
            prepost::mask_postprocess(s, ws.low.get(), 65536, ws.plane_index.get(), planes, seg->size_.w, seg->size_.h, W, H,
                                      plane_ptrs_.get());
        } else {
            int const flip = host_group_ & 1;
            ++host_group_;
            if (mask_out_[flip].size() < plane_bytes * planes) {
                synchronize();
                mask_out_[flip].allocate(plane_bytes * (size_t)std::max(planes, std::min(max_prompts_, kHostGroup) * 3));
                mask_used_[flip] = false;
            }
            if (mask_used_[flip]) CUDA_CHECK(cudaStreamWaitEvent(s, mask_free_[flip], 0));  // its previous download is done
            prepost::mask_postprocess_contiguous(s, ws.low.get(), 65536, ws.plane_index.get(), planes, seg->size_.w,
                                                 seg->size_.h, W, H, mask_out_[flip].get());
            if (ious_out) {  // tiny, and ws.iou_sel is rewritten by the next group: keep it on the work stream
                CUDA_CHECK(cudaMemcpyAsync(ious_out + (size_t)i * n, ws.iou_sel.get(), sizeof(float) * (size_t)planes,
                                           cudaMemcpyDeviceToHost, s));
                g_d2h_bytes += sizeof(float) * (size_t)planes;
            }
            CUDA_CHECK(cudaEventRecord(mask_ready_, s));
            CUDA_CHECK(cudaStreamWaitEvent(copy_out_, mask_ready_, 0));
            for (int k = 0; k < planes; ++k)
                CUDA_CHECK(cudaMemcpyAsync(planes_out[(size_t)i * n + k], mask_out_[flip].get() + plane_bytes * k, plane_bytes,
                                           cudaMemcpyDeviceToHost, copy_out_));
            g_d2h_bytes += plane_bytes * planes;
            CUDA_CHECK(cudaEventRecord(mask_free_[flip], copy_out_));
            mask_used_[flip] = true;
        }
        i = j;
    }
    if (!on_device && !async_host) {  // the caller's host buffers are complete when the call returns
        CUDA_CHECK(cudaStreamSynchronize(s));
        CUDA_CHECK(cudaStreamSynchronize(copy_out_));
    }
}

void EnvironmentImpl::low_res_logits(SegmentationImpl& seg, dlimg_b200_Prompt const& prompt, float* logits_host, float* iou_host) {
    // decode one prompt, then read back the unselected (4, 256, 256) logits and (4) IoU predictions
    uint8_t* dummy = nullptr;
    (void)dummy;
    SegmentationImpl* segs[1] = {&seg};
    {
        std::lock_guard<std::mutex> lock(mutex_);
        bind_device();
        cudaStream_t const s = stream();
        SamModel& m = model();
        DecoderWorkspace& ws = decoder_ws();
        if (!seg.encoded()) fail("segmentation handle holds no processed image");
        if (!seg.cache_.ready) m.prepare_embedding(s, seg.emb_, seg.cache_);
        float hc[4], hl[2];
        float const scale = seg.size_.scale;
        hc[0] = float(prepost::scale_coord(prompt.x0, scale));
        hc[1] = float(prepost::scale_coord(prompt.y0, scale));
        if (prompt.kind == 0) {
            hc[2] = hc[3] = float(prepost::scale_coord(0, scale));
            hl[0] = 1.0f; hl[1] = -1.0f;
        } else {
            hc[2] = float(prepost::scale_coord(prompt.x1, scale));
            hc[3] = float(prepost::scale_coord(prompt.y1, scale));
            hl[0] = 2.0f; hl[1] = 3.0f;
        }
        CUDA_CHECK(cudaMemcpyAsync(ws.coords.get(), hc, sizeof(hc), cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaMemcpyAsync(ws.labels.get(), hl, sizeof(hl), cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaStreamSynchronize(s));  // hc / hl live on this stack frame
        m.decode(s, ws, seg.cache_, 1);
        CUDA_CHECK(cudaMemcpyAsync(logits_host, ws.low.get(), sizeof(float) * 4 * 65536, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaMemcpyAsync(iou_host, ws.iou.get(), sizeof(float) * 4, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
    }
    (void)segs;
}

// ---------------------------------------------------------------------------------------------
void EnvironmentImpl::resize_longest_side(dlimg_ImageView const& v, int max_side, uint8_t* dev_out, int* out_extent) {
    std::lock_guard<std::mutex> lock(mutex_);
    bind_device();
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
    size_t const need = (size_t)v.height * size.w * bpp;
    if (resize_scratch_.size() < need) {
        CUDA_CHECK(cudaStreamSynchronize(s));
        resize_scratch_.allocate(need);
    }
    prepost::resize_srgb(s, v.pixels, v.width, v.height, v.stride, bpp, resize_tables(v.width, v.height, size.w, size.h),
                         resize_scratch_.get(), dev_out, size.w, size.h);
}

void EnvironmentImpl::image_tensor(dlimg_ImageView const& v, float* dev_out) {
    std::lock_guard<std::mutex> lock(mutex_);
    bind_device();
    check_view(v);
    prepost::image_tensor(stream(), v.pixels, v.width, v.height, v.stride, v.channels, dev_out);
}

void EnvironmentImpl::mask_postprocess(float const* dev_low_res, int count, int w, int h, uint8_t* dev_out) {
    std::lock_guard<std::mutex> lock(mutex_);
    bind_device();
    prepost::LongestSide const size = prepost::resize_longest_side(w, h, kImageSize);
    prepost::mask_postprocess_contiguous(stream(), dev_low_res, 65536, nullptr, count, size.w, size.h, w, h, dev_out);
}

void EnvironmentImpl::threshold_mask(float const* dev_logits, int th, int tw, int w, int h, uint8_t* dev_out) {
    std::lock_guard<std::mutex> lock(mutex_);
    bind_device();
    prepost::threshold_mask(stream(), dev_logits, th, tw, w, h, dev_out);
}

size_t EnvironmentImpl::encode_tap(dlimg_ImageView const* views, int count, char const* tap_name, float* dev_out, size_t capacity) {
    std::lock_guard<std::mutex> lock(mutex_);
    bind_device();
    DLIMG_ASSERT(count >= 1 && count <= max_batch_);
    for (int i = 0; i < count; ++i) check_view(views[i]);
    prepost::LongestSide const size = prepost::resize_longest_side(views[0].width, views[0].height, kImageSize);
    if (size.needs_resize && !resized_px_) resized_px_.allocate(kResizedSlot * (size_t)max_batch_);
    std::vector<enc::ImageDesc> descs((size_t)count);
    for (int i = 0; i < count; ++i) prepare_input(views[i], views[i].pixels, views[i].stride, size, i, descs[(size_t)i]);
    DeviceBuffer<float> emb((size_t)count * dec::kImgTokens * kEmbedDim);
    Tap tap;
    tap.name = tap_name;
    tap.out = dev_out;
    tap.capacity = capacity;
    encode_chunk(descs.data(), count, size, views[0].channels, emb.get(), &tap);
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

// The encoder leaves an NCHW copy next to the token-major embedding; the device->host transfer runs on the copy-out
// stream as soon as that chunk is ready, so it overlaps whatever the work stream does next.
void SegmentationImpl::embedding_nchw_async(float* out_host) {
    std::lock_guard<std::mutex> lock(env_.mutex());
    env_.bind_device();
    if (!encoded()) fail("segmentation handle holds no processed image");
    size_t const n = (size_t)dec::kImgTokens * kEmbedDim;
    CUDA_CHECK(cudaStreamWaitEvent(env_.copy_out_, emb_store_->ready(), 0));
    CUDA_CHECK(cudaMemcpyAsync(out_host, emb_nchw_, n * sizeof(float), cudaMemcpyDeviceToHost, env_.copy_out_));
    CUDA_CHECK(cudaEventRecord(emb_store_->last_read(), env_.copy_out_));
    emb_store_->mark_read();
    g_d2h_bytes += n * sizeof(float);
}

void SegmentationImpl::embedding_nchw(float* out_host) {
    embedding_nchw_async(out_host);
    CUDA_CHECK(cudaStreamSynchronize(env_.copy_out_));
}

}  // namespace dlimg
