/* ORACLE (test infrastructure only) -- plain-C restatement of the CPU-side arithmetic dlimgedit
 * performs around the two onnxruntime calls of the segmentation hot path.
 *
 *   ref_resize_longest_side   reference src/segmentation.cpp:26,60-70   (pinned: test_segmentation.cpp:15-46)
 *   ref_transform_coord       reference src/segmentation.cpp:26,72-74   (pinned: test_segmentation.cpp:48-57)
 *   ref_create_image_tensor   reference src/segmentation.cpp:81-106 + src/tensor.cpp:9-11
 *                                                                      (pinned: test_segmentation.cpp:59-83)
 *   ref_write_mask_image      reference src/segmentation.cpp:108-116   (pinned: test_segmentation.cpp:85-99)
 *   ref_resize_srgb           reference src/image.cpp:37-51 -> stb_image_resize.h v0.97 @5736b15
 *                             (depend/stb/CMakeLists.txt:6; third-party, NOT vendored in /root/reference).
 *                             Restated from the published algorithm.  PARITY UNPINNED
 *                             beyond the weak KAT test_image.cpp:51-69: stb's linear->sRGB8 step is a
 *                             104-entry table approximation (max error 0.544 ulp) that cannot be
 *                             regenerated offline; this file uses the correctly-rounded sRGB OETF, so
 *                             isolated +-1 LSB differences against real stb output are expected.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * this library.  The product never links it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- a1: ResizeLongestSide --------------------------------------------------------------- */
static int scale_coord(int coord, float scale) { return (int)(coord * scale + 0.5f); }

/* returns 1 when a resize is required (scale != 1), 0 when the original view is passed through */
int ref_resize_longest_side(int w, int h, int max_side, int* out_w, int* out_h, float* out_scale) {
    float scale = (float)max_side / (float)(w > h ? w : h);
    *out_scale = scale;
    if (scale != 1) {
        *out_w = scale_coord(w, scale);
        *out_h = scale_coord(h, scale);
        return 1;
    }
    *out_w = w;
    *out_h = h;
    return 0;
}

int ref_transform_coord(int coord, float scale) { return scale_coord(coord, scale); }

/* ---- a3: create_image_tensor ------------------------------------------------------------- */
/* channels carries the enum VALUE: mask=1 rgb=3 rgba=4 bgra=5 argb=6 (dlimgedit.hpp:29).  Like the
 * reference (tensor.cpp:9-11) rows are assumed packed: the stride field is ignored. */
void ref_create_image_tensor(const uint8_t* px, int w, int h, int channels, float* out /* h*w*3 */) {
    int cmap[3] = {0, 1, 2};
    int bpp = channels > 4 ? 4 : channels;
    if (channels == 1) { cmap[0] = cmap[1] = cmap[2] = 0; }
    else if (channels == 5) { cmap[0] = 2; cmap[1] = 1; cmap[2] = 0; }
    else if (channels == 6) { cmap[0] = 1; cmap[1] = 2; cmap[2] = 3; }
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j)
            for (int k = 0; k < 3; ++k)
                out[((size_t)i * w + j) * 3 + k] = (float)px[((size_t)i * w + j) * bpp + cmap[k]];
}

/* ---- a8: write_mask_image ---------------------------------------------------------------- */
/* logits is a (1, n, th, tw) row-major tensor; the extent (w, h) may be smaller than (tw, th): the
 * row stride comes from the tensor (KAT: tensor (1,1,2,5), extent 4x2). */
void ref_write_mask_image(const float* logits, int th, int tw, int index, int w, int h, uint8_t* out) {
    const float* plane = logits + (size_t)index * th * tw;
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j)
            out[(size_t)i * w + j] = (uint8_t)(plane[(size_t)i * tw + j] > 0 ? 255 : 0);
}

/* ---- a2: stb_image_resize v0.97 restatement ---------------------------------------------- */
static float srgb_to_linear_tab[256];
static int srgb_tab_ready = 0;

static void init_srgb_tab(void) {
    if (srgb_tab_ready) return;
    for (int i = 0; i < 256; ++i) {
        double c = i / 255.0;
        double l = c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4);
        /* stb prints its 256-entry table with 6 decimals; reproduce that rounding */
        srgb_to_linear_tab[i] = (float)(floor(l * 1e6 + 0.5) / 1e6);
    }
    srgb_tab_ready = 1;
}

void ref_srgb_decode_table(float* out256) {
    init_srgb_tab();
    memcpy(out256, srgb_to_linear_tab, sizeof(srgb_to_linear_tab));
}

uint8_t ref_linear_to_srgb8(float x) {
    if (!(x > 0.0f)) return 0;
    if (x >= 1.0f) return 255;
    double v = x <= 0.0031308 ? 12.92 * (double)x : 1.055 * pow((double)x, 1.0 / 2.4) - 0.055;
    int r = (int)floor(v * 255.0 + 0.5);
    return (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
}

static float k_mitchell(float x) {
    x = fabsf(x);
    if (x < 1.0f) return (16 + x * x * (21 * x - 36)) / 18;
    if (x < 2.0f) return (32 + x * (-60 + x * (36 - 7 * x))) / 18;
    return 0.0f;
}
static float k_catmullrom(float x) {
    x = fabsf(x);
    if (x < 1.0f) return 1 - x * x * (2.5f - 1.5f * x);
    if (x < 2.0f) return 2 - x * (4 + x * (0.5f * x - 2.5f));
    return 0.0f;
}

typedef struct {
    int n0, n1;   /* inclusive range of taps */
    float* coef;  /* n1-n0+1 weights */
} contrib_t;

static int iclamp(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

/* Upsampling (scale > 1): one contributor per OUTPUT pixel, listing INPUT taps; Catmull-Rom. */
static contrib_t* make_upsample(int in_size, int out_size, float scale) {
    contrib_t* c = (contrib_t*)calloc((size_t)out_size, sizeof(contrib_t));
    float out_radius = 2.0f * scale; /* support(1/scale) * scale */
    (void)in_size;
    for (int n = 0; n < out_size; ++n) {
        float out_center = (float)n + 0.5f;
        float lo = (out_center - out_radius) / scale;
        float hi = (out_center + out_radius) / scale;
        float in_center = out_center / scale;
        int first = (int)floorf(lo + 0.5f);
        int last = (int)floorf(hi - 0.5f);
        int cnt = last - first + 1;
        float* w = (float*)calloc((size_t)(cnt > 0 ? cnt : 1), sizeof(float));
        float total = 0;
        int i;
        for (i = 0; i <= last - first; ++i) {
            float in_px_center = (float)(i + first) + 0.5f;
            w[i] = k_catmullrom(in_center - in_px_center);
            if (i == 0 && !w[i]) { ++first; --i; continue; }
            total += w[i];
        }
        float fs = 1 / total;
        for (i = 0; i <= last - first; ++i) w[i] *= fs;
        int n1 = last;
        for (i = last - first; i >= 0; --i) { if (w[i]) break; n1 = first + i - 1; }
        c[n].n0 = first; c[n].n1 = n1; c[n].coef = w;
    }
    return c;
}

/* Downsampling (scale <= 1): one contributor per INPUT pixel (incl. clamped margin), listing OUTPUT
 * pixels it feeds; Mitchell evaluated in output space; then per-output normalisation. */
static contrib_t* make_downsample(int in_size, int out_size, float scale, int* out_margin) {
    int pixel_width = (int)ceil(2.0f * 2 / scale);
    int margin = pixel_width / 2;
    int nc = in_size + 2 * margin;
    contrib_t* c = (contrib_t*)calloc((size_t)nc, sizeof(contrib_t));
    float in_radius = 2.0f / scale;
    for (int j = 0; j < nc; ++j) {
        int n = j - margin;
        float in_center = (float)n + 0.5f;
        float lo = (in_center - in_radius) * scale;
        float hi = (in_center + in_radius) * scale;
        float out_center_of_in = in_center * scale;
        int first = (int)floorf(lo + 0.5f);
        int last = (int)floorf(hi - 0.5f);
        int cnt = last - first + 1;
        float* w = (float*)calloc((size_t)(cnt > 0 ? cnt : 1), sizeof(float));
        for (int i = 0; i <= last - first; ++i) {
            float out_px_center = (float)(i + first) + 0.5f;
            w[i] = k_mitchell(out_px_center - out_center_of_in) * scale;
        }
        int n1 = last;
        for (int i = last - first; i >= 0; --i) { if (w[i]) break; n1 = first + i - 1; }
        c[j].n0 = first; c[j].n1 = n1; c[j].coef = w;
    }
    /* normalise so each output pixel's weights sum to one (ascending contributor order) */
    for (int o = 0; o < out_size; ++o) {
        float total = 0;
        for (int j = 0; j < nc; ++j) {
            if (o >= c[j].n0 && o <= c[j].n1) total += c[j].coef[o - c[j].n0];
            else if (o < c[j].n0) break;
        }
        float s = 1 / total;
        for (int j = 0; j < nc; ++j) {
            if (o >= c[j].n0 && o <= c[j].n1) c[j].coef[o - c[j].n0] *= s;
            else if (o < c[j].n0) break;
        }
    }
    *out_margin = margin;
    return c;
}

static void free_contrib(contrib_t* c, int n) {
    for (int i = 0; i < n; ++i) free(c[i].coef);
    free(c);
}

/* horizontally resample one decoded scanline (float, linear) into out_row[out_w*ch] */
static void hresample(const uint8_t* row, int in_w, int ch, int out_w, float scale, const contrib_t* hc,
                      int hmargin, float* out_row) {
    memset(out_row, 0, sizeof(float) * (size_t)out_w * ch);
    if (scale > 1) {
        for (int x = 0; x < out_w; ++x)
            for (int k = hc[x].n0; k <= hc[x].n1; ++k) {
                int src = iclamp(k, 0, in_w - 1);
                float cf = hc[x].coef[k - hc[x].n0];
                for (int c = 0; c < ch; ++c) out_row[x * ch + c] += srgb_to_linear_tab[row[src * ch + c]] * cf;
            }
    } else {
        for (int j = 0; j < in_w + 2 * hmargin; ++j) {
            int src = iclamp(j - hmargin, 0, in_w - 1);
            int n0 = hc[j].n0 < 0 ? 0 : hc[j].n0;
            int n1 = hc[j].n1 > out_w - 1 ? out_w - 1 : hc[j].n1;
            for (int k = n0; k <= n1; ++k) {
                float cf = hc[j].coef[k - hc[j].n0];
                for (int c = 0; c < ch; ++c) out_row[k * ch + c] += srgb_to_linear_tab[row[src * ch + c]] * cf;
            }
        }
    }
}

/* u8 (in_h, in_w, ch) with byte stride -> packed u8 (out_h, out_w, ch).  All channels (alpha too) go
 * through the sRGB transfer: STBIR_ALPHA_CHANNEL_NONE, flags 0 (image.cpp:41-45). */
int ref_resize_srgb(const uint8_t* in, int in_w, int in_h, int stride, int ch, uint8_t* out, int out_w, int out_h) {
    init_srgb_tab();
    if (stride == 0) stride = in_w * ch;
    float hs = (float)out_w / in_w, vs = (float)out_h / in_h;
    int hmargin = 0, vmargin = 0;
    contrib_t* hc = hs > 1 ? make_upsample(in_w, out_w, hs) : make_downsample(in_w, out_w, hs, &hmargin);
    contrib_t* vc = vs > 1 ? make_upsample(in_h, out_h, vs) : make_downsample(in_h, out_h, vs, &vmargin);
    size_t rowf = (size_t)out_w * ch;
    float* acc = (float*)calloc((size_t)out_h * rowf, sizeof(float));
    float* hrow = (float*)malloc(sizeof(float) * rowf);
    if (vs > 1) {
        /* horizontal pass of every input row once, then gather per output row (ascending taps) */
        float* hall = (float*)malloc(sizeof(float) * rowf * (size_t)in_h);
        for (int y = 0; y < in_h; ++y) hresample(in + (size_t)y * stride, in_w, ch, out_w, hs, hc, hmargin, hall + (size_t)y * rowf);
        for (int y = 0; y < out_h; ++y)
            for (int k = vc[y].n0; k <= vc[y].n1; ++k) {
                const float* src = hall + (size_t)iclamp(k, 0, in_h - 1) * rowf;
                float cf = vc[y].coef[k - vc[y].n0];
                float* dst = acc + (size_t)y * rowf;
                for (size_t x = 0; x < rowf; ++x) dst[x] += src[x] * cf;
            }
        free(hall);
    } else {
        for (int j = 0; j < in_h + 2 * vmargin; ++j) {
            int n0 = vc[j].n0 < 0 ? 0 : vc[j].n0;
            int n1 = vc[j].n1 > out_h - 1 ? out_h - 1 : vc[j].n1;
            if (n1 < n0) continue;
            int src_y = iclamp(j - vmargin, 0, in_h - 1);
            hresample(in + (size_t)src_y * stride, in_w, ch, out_w, hs, hc, hmargin, hrow);
            for (int k = n0; k <= n1; ++k) {
                float cf = vc[j].coef[k - vc[j].n0];
                float* dst = acc + (size_t)k * rowf;
                for (size_t x = 0; x < rowf; ++x) dst[x] += hrow[x] * cf;
            }
        }
    }
    for (size_t i = 0; i < (size_t)out_h * rowf; ++i) out[i] = ref_linear_to_srgb8(acc[i]);
    free(acc); free(hrow);
    free_contrib(hc, hs > 1 ? out_w : in_w + 2 * hmargin);
    free_contrib(vc, vs > 1 ? out_h : in_h + 2 * vmargin);
    return 1;
}

/* Dump the per-axis resampling weights as a dense gather table (used by tests to check the CUDA
 * weight builder): for every output pixel, first clamped-tap index and up to max_taps weights in
 * ascending UNCLAMPED input order.  Returns the tap count per output, or -1 if max_taps too small. */
int ref_resize_weights(int in_size, int out_size, int max_taps, int* first_tap, float* weights) {
    float s = (float)out_size / in_size;
    int used = 0;
    if (s > 1) {
        contrib_t* c = make_upsample(in_size, out_size, s);
        for (int o = 0; o < out_size; ++o) {
            int cnt = c[o].n1 - c[o].n0 + 1;
            if (cnt > max_taps) { free_contrib(c, out_size); return -1; }
            if (cnt > used) used = cnt;
            first_tap[o] = c[o].n0;
            for (int i = 0; i < max_taps; ++i) weights[(size_t)o * max_taps + i] = i < cnt ? c[o].coef[i] : 0.0f;
        }
        free_contrib(c, out_size);
    } else {
        int margin = 0;
        contrib_t* c = make_downsample(in_size, out_size, s, &margin);
        int nc = in_size + 2 * margin;
        for (int o = 0; o < out_size; ++o) {
            int first = 1 << 30, cnt = 0;
            for (int i = 0; i < max_taps; ++i) weights[(size_t)o * max_taps + i] = 0.0f;
            for (int j = 0; j < nc; ++j) {
                if (o >= c[j].n0 && o <= c[j].n1) {
                    if (first == (1 << 30)) first = j - margin;
                    int idx = (j - margin) - first;
                    if (idx >= max_taps) { free_contrib(c, nc); return -1; }
                    weights[(size_t)o * max_taps + idx] = c[j].coef[o - c[j].n0];
                    cnt = idx + 1;
                }
            }
            for (int i = cnt; i < max_taps; ++i) weights[(size_t)o * max_taps + i] = 0.0f;
            first_tap[o] = first;
            if (cnt > used) used = cnt;
        }
        free_contrib(c, nc);
    }
    return used;
}
