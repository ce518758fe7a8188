// model.cu -- see model.hpp.
#include "model.hpp"

#include "kernels/gemm.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>

namespace dlimg {

// ---------------------------------------------------------------------------------------------
// TinyViT-5M as configured for MobileSAM
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kDims[4] = {64, 128, 160, 320};
constexpr int kDepths[4] = {2, 2, 6, 2};
constexpr int kHeadsCfg[4] = {2, 4, 5, 10};
constexpr int kWindows[4] = {7, 7, 14, 7};
constexpr int kRes[4] = {256, 128, 64, 64};
constexpr float kBnEps = 1e-5f;

std::vector<act_t> to_act(std::vector<float> const& v) {
    std::vector<act_t> o(v.size());
    for (size_t i = 0; i < v.size(); ++i) o[i] = f2act(v[i]);
    return o;
}

// Conv2d_BN folding: y = conv(x) * scale + shift, scale = gamma / sqrt(var + eps), shift = beta - mean * scale
void bn_fold(WeightFile const& wf, std::string const& p, int c, std::vector<float>& scale, std::vector<float>& shift) {
    auto const& g = wf.get(p + ".bn.weight", {c});
    auto const& b = wf.get(p + ".bn.bias", {c});
    auto const& m = wf.get(p + ".bn.running_mean", {c});
    auto const& v = wf.get(p + ".bn.running_var", {c});
    scale.resize((size_t)c);
    shift.resize((size_t)c);
    for (int i = 0; i < c; ++i) {
        scale[(size_t)i] = g.data[(size_t)i] / std::sqrt(v.data[(size_t)i] + kBnEps);
        shift[(size_t)i] = b.data[(size_t)i] - m.data[(size_t)i] * scale[(size_t)i];
    }
}

// k x k Conv2d_BN (cout, cin, k, k) -> GEMM operand (cout, k*k*cin), K ordered (ky, kx, ci)
Linear16 load_conv_bn(WeightFile const& wf, std::string const& p, int cout, int cin, int ks) {
    auto const& w = wf.get(p + ".c.weight", {cout, cin, ks, ks});
    std::vector<float> scale, shift;
    bn_fold(wf, p, cout, scale, shift);
    int const K = ks * ks * cin;
    std::vector<float> o((size_t)cout * K);
    for (int oc = 0; oc < cout; ++oc)
        for (int ci = 0; ci < cin; ++ci)
            for (int t = 0; t < ks * ks; ++t)
                o[(size_t)oc * K + (size_t)t * cin + ci] = w.data[((size_t)oc * cin + ci) * ks * ks + t] * scale[(size_t)oc];
    Linear16 l;
    l.n = cout;
    l.k = K;
    l.w.upload(to_act(o));
    l.b.upload(shift);
    return l;
}

DwConv load_dw_bn(WeightFile const& wf, std::string const& p, int c) {
    auto const& w = wf.get(p + ".c.weight", {c, 1, 3, 3});
    std::vector<float> scale, shift;
    bn_fold(wf, p, c, scale, shift);
    std::vector<float> o((size_t)9 * c);
    for (int ch = 0; ch < c; ++ch)
        for (int t = 0; t < 9; ++t) o[(size_t)t * c + ch] = w.data[(size_t)ch * 9 + t] * scale[(size_t)ch];
    DwConv d;
    d.c = c;
    d.w.upload(o);
    d.w16.upload(to_act(o));
    d.b.upload(shift);
    return d;
}

Linear16 load_linear16(WeightFile const& wf, std::string const& p, int n, int k, bool bias = true) {
    Linear16 l;
    l.n = n;
    l.k = k;
    l.w.upload(to_act(wf.get(p + ".weight", {n, k}).data));
    if (bias) l.b.upload(wf.get(p + ".bias", {n}).data);
    return l;
}

// Linear preceded by a LayerNorm, with the LayerNorm folded in: W'' = W * gamma (per input feature) with each row
// centred, b' = b + W beta.  The GEMM on the raw rows then yields sum_k (x_k - mean) (W gamma)_nk directly and the
// epilogue only scales by 1/std (gemm.cuh, Epilogue::ln_stats).
Linear16 load_linear16_ln(WeightFile const& wf, std::string const& p, std::string const& norm, int n, int k,
                          std::vector<float>* bias_out = nullptr) {
    auto const& w = wf.get(p + ".weight", {n, k}).data;
    auto const& b = wf.get(p + ".bias", {n}).data;
    auto const& g = wf.get(norm + ".weight", {k}).data;
    auto const& beta = wf.get(norm + ".bias", {k}).data;
    std::vector<float> wf32((size_t)n * k), bias((size_t)n);
    for (int i = 0; i < n; ++i) {
        double acc = b[(size_t)i], row_sum = 0;
        for (int j = 0; j < k; ++j) {
            double const v = (double)w[(size_t)i * k + j] * g[(size_t)j];
            row_sum += v;
            acc += (double)w[(size_t)i * k + j] * beta[(size_t)j];
        }
        double const row_mean = row_sum / k;
        for (int j = 0; j < k; ++j) wf32[(size_t)i * k + j] = (float)((double)w[(size_t)i * k + j] * g[(size_t)j] - row_mean);
        bias[(size_t)i] = (float)acc;
    }
    // The centring has to survive the rounding to 16 bits: the epilogue relies on sum_k W''_nk == 0 to cancel the row
    // mean, and a residual r_n = sum_k round16(W''_nk) (~5e-4 for K = 160) would leak mean_m * r_n * rstd_m into the output
    // -- 3 % of a unit-variance output for a token whose mean is 60 standard deviations.  Fold r_n back into the
    // smallest element of the row (finest spacing; the change, <= 5e-4 on ONE weight, is multiplied by a deviation from
    // the mean, not by the mean), twice: the residual drops to ~1e-7.
    std::vector<act_t> w16 = to_act(wf32);
    for (int i = 0; i < n; ++i) {
        act_t* row = w16.data() + (size_t)i * k;
        for (int pass = 0; pass < 2; ++pass) {
            double r = 0;
            int jmin = 0;
            for (int j = 0; j < k; ++j) {
                r += (double)act2f(row[j]);
                if (std::fabs(act2f(row[j])) < std::fabs(act2f(row[jmin]))) jmin = j;
            }
            row[jmin] = f2act((float)((double)act2f(row[jmin]) - r));
        }
    }
    Linear16 l;
    l.n = n;
    l.k = k;
    l.w.upload(w16);
    l.b.upload(bias);
    l.ln_folded = true;
    if (bias_out) *bias_out = bias;
    return l;
}

Linear32 load_linear32(WeightFile const& wf, std::string const& p, int n, int k) {
    Linear32 l;
    l.n = n;
    l.k = k;
    l.w.upload(wf.get(p + ".weight", {n, k}).data);
    l.b.upload(wf.get(p + ".bias", {n}).data);
    return l;
}

Norm load_norm(WeightFile const& wf, std::string const& p, int c) {
    Norm n;
    n.g.upload(wf.get(p + ".weight", {c}).data);
    n.b.upload(wf.get(p + ".bias", {c}).data);
    return n;
}

Linear32T load_linear32t(WeightFile const& wf, std::string const& p, int n, int k) {
    auto const& w = wf.get(p + ".weight", {n, k}).data;
    std::vector<float> t((size_t)n * k);  // [k / 4][n][4]: see decoder_tokens.cu cta_proj
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) t[((size_t)(j / 4) * n + i) * 4 + (j % 4)] = w[(size_t)i * k + j];
    Linear32T l;
    l.n = n;
    l.k = k;
    l.wt.upload(t);
    l.b.upload(wf.get(p + ".bias", {n}).data);
    return l;
}

AttnW load_attn(WeightFile const& wf, std::string const& p, int dim, int internal) {
    AttnW a;
    a.q = load_linear32t(wf, p + ".q_proj", internal, dim);
    a.k = load_linear32t(wf, p + ".k_proj", internal, dim);
    a.v = load_linear32t(wf, p + ".v_proj", internal, dim);
    a.o = load_linear32t(wf, p + ".out_proj", dim, internal);
    return a;
}

// attention_bias_idxs (a non-persistent buffer upstream): first-seen order of (|dy|, |dx|) offsets over
// all ordered pairs of window positions -> dense (heads, n, n) bias table.
std::vector<float> dense_attention_bias(HostTensor const& biases, int heads, int ws) {
    int const n = ws * ws;
    std::map<std::pair<int, int>, int> offsets;
    std::vector<int> idx((size_t)n * n);
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            std::pair<int, int> const o{std::abs(a / ws - b / ws), std::abs(a % ws - b % ws)};
            auto it = offsets.find(o);
            if (it == offsets.end()) it = offsets.emplace(o, (int)offsets.size()).first;
            idx[(size_t)a * n + b] = it->second;
        }
    int const n_off = (int)offsets.size();
    DLIMG_ASSERT(biases.shape.size() == 2 && biases.dim(0) == heads && biases.dim(1) == n_off);
    std::vector<float> dense((size_t)heads * n * n);
    for (int h = 0; h < heads; ++h)
        for (size_t i = 0; i < (size_t)n * n; ++i) dense[(size_t)h * n * n + i] = biases.data[(size_t)h * n_off + idx[i]];
    return dense;
}

void tap_act(cudaStream_t s, Tap* tap, char const* name, act_t const* p, size_t n) {
    if (!tap || !tap->name || std::strcmp(tap->name, name) != 0) return;
    if (n > tap->capacity) fail(std::string("tap buffer too small for ") + name);
    enc::act_to_f32(s, p, (int64_t)n, tap->out);
    tap->written = n;
}

}  // namespace

StageCfg SamModel::stage(int i) { return StageCfg{kDims[i], kRes[i], kDepths[i], kHeadsCfg[i], kWindows[i]}; }

// ---------------------------------------------------------------------------------------------
SamModel::SamModel(std::string const& weight_path, int num_sms) : num_sms_(num_sms) {
    WeightFile const wf = WeightFile::load(weight_path);
    std::string const E = "image_encoder.";

    // --- PatchEmbed
    {
        auto const& w = wf.get(E + "patch_embed.seq.0.c.weight", {32, 3, 3, 3});
        std::vector<float> scale, shift;
        bn_fold(wf, E + "patch_embed.seq.0", 32, scale, shift);
        std::vector<float> o(27 * 32);
        for (int oc = 0; oc < 32; ++oc)
            for (int ci = 0; ci < 3; ++ci)
                for (int t = 0; t < 9; ++t) o[(size_t)(t * 3 + ci) * 32 + oc] = w.data[((size_t)oc * 3 + ci) * 9 + t] * scale[(size_t)oc];
        enc_.conv1_w.upload(o);
        enc_.conv1_b.upload(shift);
        enc_.conv2 = load_conv_bn(wf, E + "patch_embed.seq.2", 64, 32, 3);
        {
            std::vector<uint32_t> frag(512);
            enc::patch_embed_w1_fragments(o.data(), frag.data());
            enc_.conv1_frag.upload(frag);
            std::vector<act_t> w2(64 * 288), w2p((size_t)64 * 320, f2act(0.f));
            CUDA_CHECK(cudaMemcpy(w2.data(), enc_.conv2.w.get(), w2.size() * sizeof(act_t), cudaMemcpyDeviceToHost));
            for (int n = 0; n < 64; ++n)
                for (int k = 0; k < 288; ++k) w2p[(size_t)n * 320 + k] = w2[(size_t)n * 288 + k];
            enc_.conv2_k320.upload(w2p);
            enc_.conv2_map = gemm::make_tensor_map(gemm::Operand{enc_.conv2_k320.get(), 64, 320, 320}, false, 64);
        }
    }
    // --- layer 0: MBConv x2
    for (int i = 0; i < 2; ++i) {
        std::string const p = E + "layers.0.blocks." + std::to_string(i);
        enc_.mb[i].conv1 = load_conv_bn(wf, p + ".conv1", 256, 64, 1);
        enc_.mb[i].conv2 = load_dw_bn(wf, p + ".conv2", 256);
        enc_.mb[i].conv3 = load_conv_bn(wf, p + ".conv3", 64, 256, 1);
    }
    // --- PatchMerging 0..2
    for (int i = 0; i < 3; ++i) {
        std::string const p = E + "layers." + std::to_string(i) + ".downsample";
        int const din = kDims[i], dout = kDims[i + 1];
        enc_.merge[i].conv1 = load_conv_bn(wf, p + ".conv1", dout, din, 1);
        enc_.merge[i].conv2 = load_dw_bn(wf, p + ".conv2", dout);
        enc_.merge[i].conv3 = load_conv_bn(wf, p + ".conv3", dout, dout, 1);
        enc_.merge[i].stride = (dout == 320 || dout == 448 || dout == 576) ? 1 : 2;
    }
    // --- TinyViT blocks
    for (int st = 1; st <= 3; ++st) {
        int const C = kDims[st], heads = kHeadsCfg[st], ws = kWindows[st];
        for (int i = 0; i < kDepths[st]; ++i) {
            std::string const p = E + "layers." + std::to_string(st) + ".blocks." + std::to_string(i);
            BlockW b;
            {
                std::vector<float> qkv_bias;
                b.qkv = load_linear16_ln(wf, p + ".attn.qkv", p + ".attn.norm", 3 * C, C, &qkv_bias);
                b.qkv_pad.upload(to_act(qkv_bias));  // x = 0 -> mean 0, acc 0 -> the folded bias
            }
            b.proj = load_linear16(wf, p + ".attn.proj", C, C);
            {
                std::vector<float> const dense = dense_attention_bias(wf.get(p + ".attn.attention_biases"), heads, ws);
                std::vector<uint16_t> frag(enc::attention_bias_fragment_count(heads, ws));
                enc::attention_bias_fragments(dense.data(), heads, ws, frag.data());
                b.attn_bias.upload(frag);
            }
            b.local_conv = load_dw_bn(wf, p + ".local_conv", C);
            b.fc1 = load_linear16_ln(wf, p + ".mlp.fc1", p + ".mlp.norm", 4 * C, C);
            b.fc2 = load_linear16(wf, p + ".mlp.fc2", C, 4 * C);
            enc_.blocks[st - 1].push_back(std::move(b));
        }
    }
    // --- neck
    {
        auto const& w1 = wf.get(E + "neck.0.weight", {256, 320, 1, 1});
        enc_.neck1.n = 256;
        enc_.neck1.k = 320;
        enc_.neck1.w.upload(to_act(w1.data));
        enc_.neck_ln1 = load_norm(wf, E + "neck.1", 256);
        auto const& w2 = wf.get(E + "neck.2.weight", {256, 256, 3, 3});
        std::vector<float> o((size_t)256 * 2304);
        for (int oc = 0; oc < 256; ++oc)
            for (int ci = 0; ci < 256; ++ci)
                for (int t = 0; t < 9; ++t) o[(size_t)oc * 2304 + (size_t)t * 256 + ci] = w2.data[((size_t)oc * 256 + ci) * 9 + t];
        enc_.neck2.n = 256;
        enc_.neck2.k = 2304;
        enc_.neck2.w.upload(to_act(o));
        enc_.neck_ln2 = load_norm(wf, E + "neck.3", 256);
    }

    // --- prompt encoder + mask decoder (fp32)
    std::string const P = "prompt_encoder.", D = "mask_decoder.";
    dec_.gaussian.upload(wf.get(P + "pe_layer.positional_encoding_gaussian_matrix", {2, 128}).data);
    {
        std::vector<float> pe(4 * 256);
        for (int i = 0; i < 4; ++i) {
            auto const& t = wf.get(P + "point_embeddings." + std::to_string(i) + ".weight", {1, 256});
            std::copy(t.data.begin(), t.data.end(), pe.begin() + i * 256);
        }
        dec_.point_embed.upload(pe);
    }
    dec_.not_a_point.upload(wf.get(P + "not_a_point_embed.weight", {1, 256}).data);
    dec_.no_mask.upload(wf.get(P + "no_mask_embed.weight", {1, 256}).data);
    dec_.iou_token.upload(wf.get(D + "iou_token.weight", {1, 256}).data);
    dec_.mask_tokens.upload(wf.get(D + "mask_tokens.weight", {4, 256}).data);
    // dense positional encoding of the 64 x 64 grid (a constant of the model): only its projections are kept
    DeviceBuffer<float> dense_pe((size_t)dec::kImgTokens * dec::kDim);
    dec::dense_pe(nullptr, dec_.gaussian.get(), dense_pe.get());
    // [Wa; Wb; Wc] (rows concatenated) as one 16-bit GEMM operand + the table pos @ [Wa^T | 0 | Wc^T] (see DecoderW)
    auto load_image_proj = [&](std::vector<std::pair<std::string, bool>> const& parts, Linear16& lin, DeviceBuffer<act_t>& pos_table) {
        int const n = 128 * (int)parts.size();
        std::vector<float> w((size_t)n * 256), wpos((size_t)n * 256, 0.f), bias((size_t)n);
        for (size_t i = 0; i < parts.size(); ++i) {
            auto const& pw = wf.get(parts[i].first + ".weight", {128, 256}).data;
            auto const& pb = wf.get(parts[i].first + ".bias", {128}).data;
            std::copy(pw.begin(), pw.end(), w.begin() + i * 128 * 256);
            std::copy(pb.begin(), pb.end(), bias.begin() + i * 128);
            if (parts[i].second) std::copy(pw.begin(), pw.end(), wpos.begin() + i * 128 * 256);  // this part sees keys + pos
        }
        lin.n = n;
        lin.k = 256;
        lin.w.upload(to_act(w));
        lin.b.upload(bias);
        DeviceBuffer<float> wpos_d, table((size_t)dec::kImgTokens * n);
        wpos_d.upload(wpos);
        gemm::Epilogue e;
        e.out_f32 = 1;
        e.ldc = n;
        gemm::launch_simt(nullptr, true, gemm::Operand{dense_pe.get(), dec::kImgTokens, 256, 256}, gemm::Operand{wpos_d.get(), n, 256, 256},
                          table.get(), e);  // fp32 CUDA-core GEMM, once at load
        pos_table.allocate((size_t)dec::kImgTokens * n);
        dec::f32_to_act(nullptr, table.get(), (int64_t)dec::kImgTokens * n, pos_table.get());
        CUDA_CHECK(cudaDeviceSynchronize());
    };
    for (int i = 0; i < 2; ++i) {
        std::string const p = D + "transformer.layers." + std::to_string(i);
        DecLayerW& l = dec_.layers[i];
        l.self_attn = load_attn(wf, p + ".self_attn", 256, 256);
        l.t2i_q = load_linear32t(wf, p + ".cross_attn_token_to_image.q_proj", 128, 256);
        l.t2i_o = load_linear32t(wf, p + ".cross_attn_token_to_image.out_proj", 256, 128);
        l.i2t_k = load_linear32t(wf, p + ".cross_attn_image_to_token.k_proj", 128, 256);
        l.i2t_v = load_linear32t(wf, p + ".cross_attn_image_to_token.v_proj", 128, 256);
        l.i2t_out = load_linear16(wf, p + ".cross_attn_image_to_token.out_proj", 256, 128);
        load_image_proj({{p + ".cross_attn_token_to_image.k_proj", true}, {p + ".cross_attn_token_to_image.v_proj", false},
                         {p + ".cross_attn_image_to_token.q_proj", true}},
                        dec_.kvq[i], dec_.pos_kvq[i]);
        l.n1 = load_norm(wf, p + ".norm1", 256);
        l.n2 = load_norm(wf, p + ".norm2", 256);
        l.n3 = load_norm(wf, p + ".norm3", 256);
        l.n4 = load_norm(wf, p + ".norm4", 256);
        l.lin1 = load_linear32(wf, p + ".mlp.lin1", 2048, 256);
        l.lin2 = load_linear32(wf, p + ".mlp.lin2", 256, 2048);
    }
    {
        std::string const p = D + "transformer.final_attn_token_to_image";
        dec_.final_q = load_linear32t(wf, p + ".q_proj", 128, 256);
        dec_.final_o = load_linear32t(wf, p + ".out_proj", 256, 128);
        load_image_proj({{p + ".k_proj", true}, {p + ".v_proj", false}}, dec_.kv_final, dec_.pos_kv_final);
    }
    dec_.norm_final = load_norm(wf, D + "transformer.norm_final_attn", 256);
    {
        // ConvTranspose2d(256, 64, 2, 2): weight (cin, cout, 2, 2) -> GEMM operand ((dy,dx,co), ci)
        auto const& w = wf.get(D + "output_upscaling.0.weight", {256, 64, 2, 2});
        auto const& b = wf.get(D + "output_upscaling.0.bias", {64});
        std::vector<float> o((size_t)256 * 256), bb(256);
        for (int ci = 0; ci < 256; ++ci)
            for (int co = 0; co < 64; ++co)
                for (int t = 0; t < 4; ++t) o[((size_t)t * 64 + co) * 256 + ci] = w.data[((size_t)ci * 64 + co) * 4 + t];
        for (int t = 0; t < 4; ++t)
            for (int co = 0; co < 64; ++co) bb[(size_t)t * 64 + co] = b.data[(size_t)co];
        dec_.up1.n = 256;
        dec_.up1.k = 256;
        dec_.up1.w.upload(to_act(o));
        dec_.up1.b.upload(bb);
        dec_.up_ln = load_norm(wf, D + "output_upscaling.1", 64);
        auto const& w2 = wf.get(D + "output_upscaling.3.weight", {64, 32, 2, 2});
        auto const& b2 = wf.get(D + "output_upscaling.3.bias", {32});
        std::vector<float> o2((size_t)128 * 64), bb2(128);
        for (int c1 = 0; c1 < 64; ++c1)
            for (int c2 = 0; c2 < 32; ++c2)
                for (int t = 0; t < 4; ++t) o2[((size_t)t * 32 + c2) * 64 + c1] = w2.data[((size_t)c1 * 32 + c2) * 4 + t];
        for (int t = 0; t < 4; ++t)
            for (int c2 = 0; c2 < 32; ++c2) bb2[(size_t)t * 32 + c2] = b2.data[(size_t)c2];
        dec_.up2.n = 128;
        dec_.up2.k = 64;
        dec_.up2.w.upload(to_act(o2));
        dec_.up2.b.upload(bb2);
    }
    for (int m = 0; m < 4; ++m) {
        std::string const p = D + "output_hypernetworks_mlps." + std::to_string(m) + ".layers.";
        dec_.hyper[m][0] = load_linear32t(wf, p + "0", 256, 256);
        dec_.hyper[m][1] = load_linear32t(wf, p + "1", 256, 256);
        dec_.hyper[m][2] = load_linear32t(wf, p + "2", 32, 256);
    }
    dec_.iou[0] = load_linear32t(wf, D + "iou_prediction_head.layers.0", 256, 256);
    dec_.iou[1] = load_linear32t(wf, D + "iou_prediction_head.layers.1", 256, 256);
    dec_.iou[2] = load_linear32t(wf, D + "iou_prediction_head.layers.2", 4, 256);

    CUDA_CHECK(cudaDeviceSynchronize());
}

// ---------------------------------------------------------------------------------------------
EncoderWorkspace::EncoderWorkspace(int mb) : max_batch(mb) {
    size_t const B = (size_t)mb;
#if DLIMG_B200_ALT  // conv1 activation and im2col matrix of the unfused PatchEmbed / neck (17 + 38 MB per image)
    c1.allocate(B * 512 * 512 * 32);
    col.allocate(B * 65536 * 288);
#endif
    xa.allocate(B * 65536 * 64);
    xb.allocate(B * 65536 * 64);
    for (auto& b : big) b.allocate(B * 65536 * 256);  // expanded MBConv / qkv / attention-output sized scratch
    stats.allocate(B * 16384);
    stats_parts.allocate(B * 16384 * 2);
}

DecoderWorkspace::DecoderWorkspace(int mp) : max_prompts(mp) {
    size_t const P = (size_t)mp;
    param_block.allocate(DecoderParams::bytes(mp));
    for (auto* b : {&tok0, &queries}) b->allocate(P * 7 * 256);
    tmp.allocate((size_t)kMlpSplit * gemm::split_rows(P * 7) * 256);  // split-K partials of the token MLP's second Linear
    t128a.allocate(P * 7 * 128);
    t128b.allocate(P * 7 * 128);
    t128c.allocate(P * 7 * 128);
    hid.allocate(P * 7 * 2048);
    hyper.allocate(P * 4 * 32);
    iou.allocate(P * 4);
    keys.allocate(P * 4096 * 256);
    big.allocate(P * 4096 * 256);
    kvq.allocate(P * 4096 * 384);
    ao.allocate(P * 4096 * 128);
    low.allocate(P * 4 * 65536);
    plane_index.allocate(P * 3);
    iou_sel.allocate(P * 3);
    t2i_scratch.allocate(P * dec::kT2iScratchPerPrompt);
}

DecoderParams DecoderWorkspace::layout(uint8_t* base, int P) {
    DecoderParams d;
    d.coords = reinterpret_cast<float*>(base);
    d.labels = d.coords + (size_t)P * 4;
    d.keys0 = reinterpret_cast<act_t const**>(d.labels + (size_t)P * 2);
    d.kvq0 = d.keys0 + P;
    return d;
}

// ---------------------------------------------------------------------------------------------
void SamModel::gemm16(cudaStream_t s, act_t const* a, int64_t rows, Linear16 const& l, void* out, int act,
                      act_t const* residual, float2 const* ln_stats, bool out_f32, int ln_parts, float2* stats_out, int res_mod) const {
    gemm::Operand A{a, rows, l.k, l.k};
    gemm::Operand B{l.w.get(), l.n, l.k, l.k};
    gemm::Epilogue e;
    e.bias = l.b.get();
    e.residual = residual;
    e.res_mod = res_mod;
    e.act = act;
    e.out_f32 = out_f32 ? 1 : 0;
    e.ldc = l.n;
    if (ln_stats) {
        DLIMG_ASSERT(l.ln_folded);
        e.ln_stats = ln_stats;
        e.ln_parts = ln_parts;
    }
    e.stats_out = stats_out;
    gemm::launch(s, false, A, B, out, e, num_sms_);
}

void SamModel::gemm32(cudaStream_t s, float const* a, int64_t rows, Linear32 const& l, float* out, int act) const {
    gemm::Operand A{a, rows, l.k, l.k};
    gemm::Operand B{l.w.get(), l.n, l.k, l.k};
    gemm::Epilogue e;
    e.bias = l.b.get();
    e.act = act;
    e.out_f32 = 1;
    e.ldc = l.n;
    gemm::launch(s, true, A, B, out, e, num_sms_);
}

void SamModel::encode(cudaStream_t s, EncoderWorkspace& ws, enc::ImageDesc const* images, int batch, int w, int h,
                      int channels, float* emb_nchw_out, act_t* keys0_out, act_t* kvq0_out, Tap* tap, bool finish) const {
    DLIMG_ASSERT(batch >= 1 && batch <= ws.max_batch);
    int64_t const B = batch;
    using gemm::ACT_GELU;
    using gemm::ACT_NONE;

    // PatchEmbed: preprocess + conv1 + GELU + conv2 in one kernel (DLIMG_B200_UNFUSED_PATCH=1 keeps the three-kernel
    // form: conv1, im2col, GEMM -- used to cross-check the fused kernel)
    act_t* x = ws.xa.get();
    act_t* y = ws.xb.get();
#if DLIMG_B200_ALT
    static bool const unfused_patch = kActBf16 || dev_switch("DLIMG_B200_UNFUSED_PATCH");
    if (unfused_patch) {
        enc::conv1_preprocess(s, images, batch, w, h, channels, enc_.conv1_w.get(), enc_.conv1_b.get(), ws.c1.get());
        tap_act(s, tap, "conv1", ws.c1.get(), (size_t)B * 512 * 512 * 32);
        enc::im2col3x3(s, ws.c1.get(), batch, 512, 512, 32, 2, ws.col.get());
        gemm16(s, ws.col.get(), B * 65536, enc_.conv2, x, ACT_NONE, nullptr);
    } else
#endif
    {
        bool const want_c1 = tap && tap->name && std::strcmp(tap->name, "conv1") == 0;
        // the (B, 512, 512, 32) conv1 activation only exists for this debug tap (eager path): 17 MB per image on demand
        if (want_c1 && !ws.c1.get()) ws.c1.allocate((size_t)ws.max_batch * 512 * 512 * 32);
        enc::patch_embed(s, images, batch, w, h, channels, enc_.conv1_frag.get(), enc_.conv1_b.get(), enc_.conv2_map,
                         enc_.conv2.b.get(), x, want_c1 ? ws.c1.get() : nullptr, num_sms_);
        tap_act(s, tap, "conv1", ws.c1.get(), (size_t)B * 512 * 512 * 32);
    }
    tap_act(s, tap, "patch_embed", x, (size_t)B * 65536 * 64);

    // layer 0: MBConv x2: 1x1 expand + GELU (GEMM), then depthwise 3x3 + GELU + 1x1 project + shortcut + GELU in one
    // kernel (DLIMG_B200_UNFUSED_MBCONV=1 keeps the depthwise kernel + project GEMM, used to cross-check)
    static bool const unfused_mbconv = kActBf16 || dev_switch("DLIMG_B200_UNFUSED_MBCONV");
    for (int i = 0; i < 2; ++i) {
        MBConvW const& m = enc_.mb[i];
        gemm16(s, x, B * 65536, m.conv1, ws.big[0].get(), ACT_GELU, nullptr);
        if (unfused_mbconv) {
            enc::dwconv3x3(s, ws.big[0].get(), batch, 256, 256, 256, 1, m.conv2.w.get(), m.conv2.w16.get(), m.conv2.b.get(), true, ws.big[1].get());
            gemm16(s, ws.big[1].get(), B * 65536, m.conv3, y, ACT_GELU, x);
        } else {
            CUtensorMap const in_map = gemm::make_tensor_map_nhwc(ws.big[0].get(), batch, 256, 256, 256, enc::mbconv_tail_unit_channels(), 18, 10);
            CUtensorMap const w3_map = gemm::make_tensor_map(gemm::Operand{m.conv3.w.get(), 64, 256, 256}, false, 64);
            enc::mbconv_tail(s, in_map, batch, m.conv2.w16.get(), m.conv2.b.get(), w3_map, m.conv3.b.get(), x, y, num_sms_);
        }
        std::swap(x, y);
        tap_act(s, tap, i == 0 ? "mb0" : "mb1", x, (size_t)B * 65536 * 64);
    }

    static bool const merge_stats = !kActBf16 && !dev_switch("DLIMG_B200_LN_STATS_KERNEL");  // A/B switch
    // conv3 of a PatchMerging block also leaves the LayerNorm row sums of its output for the qkv GEMM of the stage's
    // first block (same form as the fc2 / fused-MLP epilogues of the later blocks)
    auto merge = [&](MergeW const& m, int res, char const* name) {
        int const out_res = m.stride == 2 ? res / 2 : res;
        int const dout = m.conv1.n;
        gemm16(s, x, B * res * res, m.conv1, ws.big[0].get(), ACT_GELU, nullptr);
        enc::dwconv3x3(s, ws.big[0].get(), batch, res, res, dout, m.stride, m.conv2.w.get(), m.conv2.w16.get(), m.conv2.b.get(),
                       true, ws.big[1].get());
        gemm16(s, ws.big[1].get(), B * out_res * out_res, m.conv3, y, ACT_NONE, nullptr, nullptr, false, 0,
               merge_stats ? ws.stats_parts.get() : nullptr);
        std::swap(x, y);
        tap_act(s, tap, name, x, (size_t)B * out_res * out_res * dout);
    };
    merge(enc_.merge[0], 256, "layer0");

    for (int st = 1; st <= 3; ++st) {
        StageCfg const c = stage(st);
        int const C = c.dim, L = c.res * c.res;
        int const rows = (int)(B * L);
        float2* stats = ws.stats.get();
        int const fc2_parts = C / gemm::pick_block_n(C);  // N tiles of fc2 = partial sums per row
        static bool const fused_mlp = !dev_switch("DLIMG_B200_UNFUSED_MLP");
        static bool const lc_tma = !kActBf16 && !dev_switch("DLIMG_B200_LOCAL_CONV_REG");  // A/B switch
        for (int i = 0; i < c.depth; ++i) {
            BlockW const& b = enc_.blocks[st - 1][(size_t)i];
            std::string const tn = "s" + std::to_string(st) + "b" + std::to_string(i);
            // attention branch on the un-partitioned grid: LN statistics, QKV with the LayerNorm folded into the GEMM,
            // windowed attention (does the partition / zero padding / un-partition itself), proj + residual in place
            // (the row sums come from the epilogue of whatever produced x: PatchMerging conv3, fc2 or the fused MLP)
            if (i == 0 && !merge_stats) {
                enc::layernorm_stats(s, x, rows, C, 1e-5f, stats);
                gemm16(s, x, rows, b.qkv, ws.big[1].get(), ACT_NONE, nullptr, stats);
            } else {
                gemm16(s, x, rows, b.qkv, ws.big[1].get(), ACT_NONE, nullptr, ws.stats_parts.get(), false, fc2_parts);
            }
            tap_act(s, tap, (tn + ".qkv").c_str(), ws.big[1].get(), (size_t)rows * 3 * C);
            enc::window_attention(s, ws.big[1].get(), batch, c.res, c.ws, c.heads, b.qkv_pad.get(), b.attn_bias.get(),
                                  ws.big[2].get(), num_sms_);
            tap_act(s, tap, (tn + ".att").c_str(), ws.big[2].get(), (size_t)rows * C);
            gemm16(s, ws.big[2].get(), rows, b.proj, x, ACT_NONE, x);
            tap_act(s, tap, (tn + ".proj").c_str(), x, (size_t)rows * C);
            // local depthwise conv (no activation, no residual).  fp32 accumulation: its output IS the trunk (it replaces
            // x), and the packed-half kernel cost 0.0006 of mask IoU here for 3 % of the step -- not worth it.
            int lc_parts = 1;  // partial row sums per pixel
            if (lc_tma && enc::local_conv_tma_supported(c.res, c.res, C)) {
                enc::local_conv_tma(s, x, batch, c.res, c.res, C, b.local_conv.w.get(), b.local_conv.b.get(), y, stats, num_sms_);
                lc_parts = enc::local_conv_parts(C);
            } else {
                enc::dwconv3x3_stats(s, x, batch, c.res, c.res, C, b.local_conv.w.get(), b.local_conv.b.get(), y, stats);
            }
            tap_act(s, tap, (tn + ".lc").c_str(), y, (size_t)rows * C);
            // MLP branch: LN folded into fc1 (row sums from the depthwise kernel above) + GELU, fc2 + residual
            float2* const next_stats = i + 1 < c.depth ? ws.stats_parts.get() : nullptr;
            if (fused_mlp && gemm::mlp_fused_supported(C)) {
                // one kernel: fc1 (+ folded LN, GELU) -> hidden activation in TMEM / shared memory -> fc2 + residual
                gemm::launch_mlp_fused(s, y, rows, C, b.fc1.w.get(), b.fc1.b.get(), stats, 1e-5f, b.fc2.w.get(), b.fc2.b.get(), y,
                                       next_stats, num_sms_);
            } else {
                gemm16(s, y, rows, b.fc1, ws.big[1].get(), ACT_GELU, nullptr, stats, false, lc_parts);
                // fc2 + residual; its epilogue also leaves the LayerNorm row sums of the result for the next block's qkv
                gemm16(s, ws.big[1].get(), rows, b.fc2, y, ACT_NONE, y, nullptr, false, 0, next_stats);
            }
            std::swap(x, y);
            tap_act(s, tap, tn.c_str(), x, (size_t)rows * C);
        }
        if (st < 3) merge(enc_.merge[st], c.res, st == 1 ? "layer1" : "layer2");
    }
    tap_act(s, tap, "layer3", x, (size_t)B * 4096 * 320);

    // neck: 1x1 conv -> LayerNorm2d -> 3x3 conv -> LayerNorm2d (token-major, so LayerNorm2d is a row LayerNorm)
    gemm16(s, x, B * 4096, enc_.neck1, ws.big[0].get(), ACT_NONE, nullptr);
    enc::layernorm_rows(s, ws.big[0].get(), (int)(B * 4096), 256, nullptr, enc_.neck_ln1.g.get(), enc_.neck_ln1.b.get(), 1e-6f,
                        ws.big[1].get(), false);
    tap_act(s, tap, "neck1", ws.big[1].get(), (size_t)B * 4096 * 256);
#if DLIMG_B200_ALT
    static bool const neck_im2col = kActBf16 || dev_switch("DLIMG_B200_NECK_IM2COL");  // A/B switch
    if (neck_im2col) {
        enc::im2col3x3(s, ws.big[1].get(), batch, 64, 64, 256, 1, ws.col.get());
        gemm16(s, ws.col.get(), B * 4096, enc_.neck2, ws.big[0].get(), ACT_NONE, nullptr);
    } else
#endif
    {  // implicit GEMM: shifted 4D TMA boxes of the NHWC activation are the A operand
        gemm::Epilogue e;
        e.bias = enc_.neck2.b.get();
        e.ldc = enc_.neck2.n;
        gemm::launch_conv3x3(s, ws.big[1].get(), batch, 64, 64, 256, gemm::Operand{enc_.neck2.w.get(), enc_.neck2.n, enc_.neck2.k, enc_.neck2.k},
                             ws.big[0].get(), e, num_sms_);
    }
    if (finish) neck_finish(s, ws, batch, emb_nchw_out, keys0_out, kvq0_out);
}

// Final LayerNorm2d of the neck -> the reference's NCHW fp32 embedding + the decoder's 16-bit layer-0 image stream, then the
// layer-0 image-side projections of the two-way transformer (tokens -> image keys / values, image -> tokens queries): they
// depend on the image only, so they are computed here, once per image and batched over the chunk, not per prompt.
void SamModel::neck_finish(cudaStream_t s, EncoderWorkspace& ws, int batch, float* emb_nchw_out, act_t* keys0_out, act_t* kvq0_out) const {
    enc::layernorm256_tokens_nchw(s, ws.big[0].get(), batch, 4096, enc_.neck_ln2.g.get(), enc_.neck_ln2.b.get(), 1e-6f,
                                  dec_.no_mask.get(), keys0_out, emb_nchw_out);
    gemm16(s, keys0_out, (int64_t)batch * dec::kImgTokens, dec_.kvq[0], kvq0_out, gemm::ACT_NONE, dec_.pos_kvq[0].get(), nullptr, false,
           0, nullptr, dec::kImgTokens);
}

// ---------------------------------------------------------------------------------------------
void SamModel::lin(cudaStream_t s, float const* x, int64_t xs, float const* x2, int rows, Linear32 const& l, bool relu,
                   float* y, int64_t ys) const {
    // token-side Linears on contiguous rows go to the tensor cores (tf32);
    // the strided / position-encoded / tiny ones stay on the CUDA-core kernel
    if (!x2 && xs == l.k && ys == l.n && l.n % 16 == 0) {  // by shape only: a prompt gets the same bits alone or in a batch
        gemm32(s, x, rows, l, y, relu ? gemm::ACT_RELU : gemm::ACT_NONE);
        return;
    }
    dec::linear_small(s, x, xs, x2, xs, rows, l.k, l.w.get(), l.b.get(), l.n, relu, y, ys);
}

void SamModel::decode(cudaStream_t s, DecoderWorkspace& ws, int P, int mask_mode) const {
    DLIMG_ASSERT(P >= 1 && P <= ws.max_prompts);
    DecoderParams const prm = DecoderWorkspace::layout(ws.param_block.get(), P);
    int const R = P * dec::kTokens;
    int64_t const IR = (int64_t)P * dec::kImgTokens;  // image-side rows
    int64_t const kvq_stride = (int64_t)dec::kImgTokens * 384, kv_stride = (int64_t)dec::kImgTokens * 256;

    dec::PromptParams pp{dec_.gaussian.get(), dec_.point_embed.get(), dec_.not_a_point.get(), dec_.iou_token.get(),
                         dec_.mask_tokens.get()};
    dec::prompt_tokens(s, prm.coords, prm.labels, P, pp, ws.tok0.get(), ws.queries.get());

    for (int li = 0; li < 2; ++li) {
        DecLayerW const& l = dec_.layers[li];
        bool const first = li == 0;
        // (1) token self-attention (layer 0: no positional encoding, output replaces the queries) + LayerNorm + the
        // query projection of (2), one kernel
        {
            dec::TokenAttnBlock a;
            a.queries = ws.queries.get();
            a.pe = ws.tok0.get();
            a.wq_t = l.self_attn.q.wt.get(); a.bq = l.self_attn.q.b.get();
            a.wk_t = l.self_attn.k.wt.get(); a.bk = l.self_attn.k.b.get();
            a.wv_t = l.self_attn.v.wt.get(); a.bv = l.self_attn.v.b.get();
            a.wo_t = l.self_attn.o.wt.get(); a.bo = l.self_attn.o.b.get();
            a.gamma = l.n1.g.get(); a.beta = l.n1.b.get();
            a.with_pe = first ? 0 : 1;
            a.residual = first ? 0 : 1;
            a.w_next_t = l.t2i_q.wt.get(); a.b_next = l.t2i_q.b.get();
            a.out_next = ws.t128a.get();
            dec::token_attn_block(s, a, P);
        }
        // image-side projections [K | V | Q] of this layer from the image stream: layer 0's depend on the image only and
        // come with the embedding (per-prompt tables); layer 1's are one GEMM with the position terms added by its epilogue
        if (!first)
            gemm16(s, ws.keys.get(), IR, dec_.kvq[1], ws.kvq.get(), gemm::ACT_NONE, dec_.pos_kvq[1].get(), nullptr, false, 0, nullptr,
                   dec::kImgTokens);
        // (2) tokens attend to the image (split-key partials), then merge + out projection + residual + LayerNorm
        dec::token_to_image_attention(s, ws.t128a.get(), ws.kvq.get(), first ? prm.kvq0 : nullptr, kvq_stride, 384, 128, P,
                                      ws.t2i_scratch.get(), nullptr);
        {
            dec::TokenPostT2i t{ws.t2i_scratch.get(), ws.queries.get(), l.t2i_o.wt.get(), l.t2i_o.b.get(), l.n2.g.get(), l.n2.b.get()};
            dec::token_post_t2i(s, t, P);
        }
        // (3) token MLP on the tensor cores (7 rows per prompt, 4 MB of weights read once per pass)
        lin(s, ws.queries.get(), 256, nullptr, R, l.lin1, true, ws.hid.get(), 2048);
        {   // second Linear (K = 2048) split over k: 16 output tiles with a 64-step k-loop each left 130 SMs idle for 22 us
            gemm::Epilogue e;
            e.out_f32 = 1;
            e.ldc = 256;
            e.ksplit = kMlpSplit;
            gemm::launch(s, true, gemm::Operand{ws.hid.get(), R, 2048, 2048}, gemm::Operand{l.lin2.w.get(), 256, 2048, 2048}, ws.tmp.get(), e,
                         num_sms_);
        }
        // residual + LayerNorm + the token-side projections of (4) (and, after the last layer, of the final attention)
        {
            dec::TokenPostMlp m;
            m.queries = ws.queries.get();
            m.mlp_out = ws.tmp.get();
            m.mlp_parts = kMlpSplit;
            m.mlp_part_stride = gemm::split_rows(R) * 256;
            m.mlp_bias = l.lin2.b.get();
            m.pe = ws.tok0.get();
            m.gamma = l.n3.g.get(); m.beta = l.n3.b.get();
            m.count = first ? 2 : 3;
            m.w_t[0] = l.i2t_k.wt.get(); m.b[0] = l.i2t_k.b.get(); m.with_pe[0] = 1; m.out[0] = ws.t128a.get();
            m.w_t[1] = l.i2t_v.wt.get(); m.b[1] = l.i2t_v.b.get(); m.with_pe[1] = 0; m.out[1] = ws.t128b.get();
            m.w_t[2] = dec_.final_q.wt.get(); m.b[2] = dec_.final_q.b.get(); m.with_pe[2] = 1; m.out[2] = ws.t128c.get();
            dec::token_post_mlp(s, m, P);
        }
        // (4) image attends to the tokens: keys <- LN(keys + out_proj(attn)); Q is column block 2 of [K | V | Q]
        dec::image_to_token_attention_mma(s, ws.kvq.get(), first ? prm.kvq0 : nullptr, kvq_stride, 384, 256, ws.t128a.get(),
                                          ws.t128b.get(), P, ws.ao.get());
        {   // out projection + residual + LayerNorm in one GEMM (Epilogue::fuse = 3).  Layer 0 reads its residual, the image's
            // own prompt-independent keys, through the per-prompt table and writes ws.keys; layer 1 updates ws.keys in place
            // (a tile reads its residual rows into shared memory before it writes the same rows).
            gemm::Epilogue e;
            e.bias = l.i2t_out.b.get();
            e.ldc = 256;
            e.fuse = 3;
            e.fuse_a = l.n4.g.get();
            e.fuse_b = l.n4.b.get();
            e.ln_eps = 1e-5f;
            if (first) {
                e.res_table = reinterpret_cast<void const* const*>(prm.keys0);
                e.res_mod = dec::kImgTokens;
            } else {
                e.residual = ws.keys.get();
            }
            gemm::launch(s, false, gemm::Operand{ws.ao.get(), IR, 128, 128}, gemm::Operand{l.i2t_out.w.get(), 256, 128, 128}, ws.keys.get(), e,
                         num_sms_);
        }
    }
    // final token -> image attention (its query projection came out of the last token_post_mlp)
    gemm16(s, ws.keys.get(), IR, dec_.kv_final, ws.kvq.get(), gemm::ACT_NONE, dec_.pos_kv_final.get(), nullptr, false, 0, nullptr,
           dec::kImgTokens);
    dec::token_to_image_attention(s, ws.t128c.get(), ws.kvq.get(), nullptr, kv_stride, 256, 128, P, ws.t2i_scratch.get(), nullptr);
    {
        dec::TokenPostT2i t{ws.t2i_scratch.get(), ws.queries.get(), dec_.final_o.wt.get(), dec_.final_o.b.get(), dec_.norm_final.g.get(),
                            dec_.norm_final.b.get()};
        dec::token_post_t2i(s, t, P);
    }

    // IoU head on the iou token, hypernetwork MLPs on the four mask tokens: one kernel for the five 3-layer MLPs
    {
        dec::TokenMlp3 h;
        for (int l = 0; l < 3; ++l) {
            h.w[0][l] = dec_.iou[l].wt.get();
            h.b[0][l] = dec_.iou[l].b.get();
            for (int m = 0; m < 4; ++m) {
                h.w[1 + m][l] = dec_.hyper[m][l].wt.get();
                h.b[1 + m][l] = dec_.hyper[m][l].b.get();
            }
        }
        dec::token_mlp3(s, ws.queries.get(), P, h, ws.hyper.get(), ws.iou.get());
    }

    // upscaling: two transposed 2x2/stride-2 convolutions as GEMMs in a blocked pixel layout, with the LayerNorm2d + GELU
    // between them and the final hypernetwork product folded into their epilogues (gemm.cuh Epilogue::fuse): the
    // (P, 65536, 32) upscaled embedding is never written, only the (P, 4, 256, 256) logits
    {
        gemm::Epilogue e;
        e.bias = dec_.up1.b.get();
        e.act = gemm::ACT_GELU;
        e.ldc = dec_.up1.n;
        e.fuse = 1;
        e.fuse_a = dec_.up_ln.g.get();
        e.fuse_b = dec_.up_ln.b.get();
        gemm::launch(s, false, gemm::Operand{ws.keys.get(), IR, 256, 256}, gemm::Operand{dec_.up1.w.get(), 256, 256, 256}, ws.big.get(), e,
                     num_sms_);
        gemm::Epilogue e2;
        e2.bias = dec_.up2.b.get();
        e2.act = gemm::ACT_GELU;
        e2.ldc = dec_.up2.n;
        e2.fuse = 2;
        e2.fuse_a = ws.hyper.get();
        e2.fuse_b = ws.iou.get();
        e2.fuse_mode = mask_mode;
        e2.fuse_out = ws.low.get();
        gemm::launch(s, false, gemm::Operand{ws.big.get(), IR * 4, 64, 64}, gemm::Operand{dec_.up2.w.get(), 128, 64, 64}, nullptr, e2, num_sms_);
    }
}

}  // namespace dlimg
