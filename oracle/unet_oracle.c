/*
 * Plain-C restatement of the hot path, independent of PyTorch -- TEST INFRASTRUCTURE ONLY.
 * Nothing under inconsistencymasks_b200/ links or loads this file; tests/ use it to
 * cross-check oracle/ref_unet.py (layout mistakes show up as disagreement between two
 * independently written restatements) and oracle/ref_im.py.
 *
 * Follows the reference (paths relative to the reference repository):
 *   unet.py:4-9     input_block    x/255 -> 1x1 conv + ReLU -> BN
 *   unet.py:11-19   encoder_block  3x3 conv + ReLU -> 1x1 conv + ReLU -> BN -> (skip) -> 2x2 max-pool
 *   unet.py:22-29   bottleneck     3x3 conv + ReLU -> 1x1 conv + ReLU -> BN
 *   unet.py:31-43   decoder_block  nearest-upsample 2x + add skip -> 1x1 conv + ReLU -> BN
 *                                  -> 3x3 conv + ReLU -> 1x1 conv + ReLU -> BN
 *   unet.py:63      out            1x1 conv -> sigmoid | softmax (float32)
 *   functions.py:3104-3120  pred_masks_to_im_binary,  functions.py:3123-3137  _multiclass
 * Keras semantics encoded: Conv2D 'same' zero padding, stride 1, kernel HWIO, bias;
 * BatchNormalization inference form with epsilon 1e-3; parity at the TensorFlow boundary is
 * UNPINNED (TensorFlow is not available, see oracle/ref_unet.py).
 * Weights arrive in Keras get_weights() order (104 arrays).  All tensors are float32 NHWC.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BN_EPS 1e-3f

typedef struct { const float *const *w; int i; } cursor_t;

static float *conv_relu(const float *x, int n, int h, int w, int cin, int cout, int ks, cursor_t *cur, int relu) {
    const float *k = cur->w[cur->i], *b = cur->w[cur->i + 1];
    cur->i += 2;
    float *y = (float *)malloc(sizeof(float) * (size_t)n * h * w * cout);
    const int r = ks / 2;
    for (int img = 0; img < n; ++img)
        for (int yy = 0; yy < h; ++yy)
            for (int xx = 0; xx < w; ++xx) {
                float *o = y + (((size_t)img * h + yy) * w + xx) * cout;
                for (int co = 0; co < cout; ++co) o[co] = b[co];
                for (int dy = 0; dy < ks; ++dy) {
                    const int sy = yy + dy - r;
                    if (sy < 0 || sy >= h) continue;
                    for (int dx = 0; dx < ks; ++dx) {
                        const int sx = xx + dx - r;
                        if (sx < 0 || sx >= w) continue;
                        const float *in = x + (((size_t)img * h + sy) * w + sx) * cin;
                        const float *kk = k + ((size_t)(dy * ks + dx) * cin) * cout;     /* HWIO */
                        for (int ci = 0; ci < cin; ++ci) {
                            const float v = in[ci];
                            const float *kr = kk + (size_t)ci * cout;
                            for (int co = 0; co < cout; ++co) o[co] += v * kr[co];
                        }
                    }
                }
                if (relu)
                    for (int co = 0; co < cout; ++co) o[co] = o[co] > 0.f ? o[co] : 0.f;
            }
    return y;
}

static void batchnorm(float *x, size_t px, int ch, cursor_t *cur) {
    const float *g = cur->w[cur->i], *be = cur->w[cur->i + 1], *mu = cur->w[cur->i + 2], *var = cur->w[cur->i + 3];
    cur->i += 4;
    for (int c = 0; c < ch; ++c) {
        const float scale = g[c] / sqrtf(var[c] + BN_EPS);
        const float shift = be[c] - mu[c] * scale;
        for (size_t p = 0; p < px; ++p) x[p * ch + c] = x[p * ch + c] * scale + shift;
    }
}

static float *maxpool(const float *x, int n, int h, int w, int ch) {
    const int ho = h / 2, wo = w / 2;
    float *y = (float *)malloc(sizeof(float) * (size_t)n * ho * wo * ch);
    for (int img = 0; img < n; ++img)
        for (int yy = 0; yy < ho; ++yy)
            for (int xx = 0; xx < wo; ++xx)
                for (int c = 0; c < ch; ++c) {
                    float m = -INFINITY;
                    for (int dy = 0; dy < 2; ++dy)
                        for (int dx = 0; dx < 2; ++dx) {
                            const float v = x[(((size_t)img * h + 2 * yy + dy) * w + 2 * xx + dx) * ch + c];
                            if (v > m) m = v;
                        }
                    y[(((size_t)img * ho + yy) * wo + xx) * ch + c] = m;
                }
    return y;
}

/* nearest-upsample 2x of lo [n,h/2,w/2,ch] added to skip [n,h,w,ch] */
static float *upsample_add(const float *lo, const float *skip, int n, int h, int w, int ch) {
    float *y = (float *)malloc(sizeof(float) * (size_t)n * h * w * ch);
    for (int img = 0; img < n; ++img)
        for (int yy = 0; yy < h; ++yy)
            for (int xx = 0; xx < w; ++xx)
                for (int c = 0; c < ch; ++c)
                    y[(((size_t)img * h + yy) * w + xx) * ch + c] =
                        lo[(((size_t)img * (h / 2) + yy / 2) * (w / 2) + xx / 2) * ch + c] +
                        skip[(((size_t)img * h + yy) * w + xx) * ch + c];
    return y;
}

/* act: 0 sigmoid, 1 softmax.  Returns 0, or -1 when the weight list length does not match. */
int oracle_unet_forward(const uint8_t *images, int n, int h, int w, int c, int k, double alpha, int ks, int act,
                        const float *const *weights, int n_weights, float *probs) {
    int f[5];
    const int base[5] = {16, 32, 64, 128, 256};
    for (int i = 0; i < 5; ++i) f[i] = (int)(base[i] * alpha);
    cursor_t cur = {weights, 0};
    const size_t px0 = (size_t)n * h * w;
    float *x = (float *)malloc(sizeof(float) * px0 * c);
    for (size_t i = 0; i < px0 * c; ++i) x[i] = (float)images[i] / 255.0f;                 /* unet.py:5 */
    float *t = conv_relu(x, n, h, w, c, f[0], 1, &cur, 1); free(x); x = t;                  /* unet.py:6 */
    batchnorm(x, px0, f[0], &cur);
    float *skips[4];
    int hh = h, ww = w, cin = f[0];
    for (int l = 0; l < 4; ++l) {                                                           /* unet.py:51-54 */
        t = conv_relu(x, n, hh, ww, cin, f[l], ks, &cur, 1); free(x); x = t;
        t = conv_relu(x, n, hh, ww, f[l], f[l], 1, &cur, 1); free(x); x = t;
        batchnorm(x, (size_t)n * hh * ww, f[l], &cur);
        skips[l] = x;
        x = maxpool(x, n, hh, ww, f[l]);
        hh /= 2; ww /= 2; cin = f[l];
    }
    t = conv_relu(x, n, hh, ww, cin, f[4], ks, &cur, 1); free(x); x = t;                    /* unet.py:56 */
    t = conv_relu(x, n, hh, ww, f[4], f[3], 1, &cur, 1); free(x); x = t;
    batchnorm(x, (size_t)n * hh * ww, f[3], &cur);
    cin = f[3];
    const int c1s[4] = {f[3], f[2], f[1], f[0]}, c2s[4] = {f[2], f[1], f[0], f[0]};
    for (int l = 0; l < 4; ++l) {                                                           /* unet.py:58-61 */
        hh *= 2; ww *= 2;
        t = upsample_add(x, skips[3 - l], n, hh, ww, cin); free(x); free(skips[3 - l]); x = t;
        t = conv_relu(x, n, hh, ww, cin, c1s[l], 1, &cur, 1); free(x); x = t;
        batchnorm(x, (size_t)n * hh * ww, c1s[l], &cur);
        t = conv_relu(x, n, hh, ww, c1s[l], c1s[l], ks, &cur, 1); free(x); x = t;
        t = conv_relu(x, n, hh, ww, c1s[l], c2s[l], 1, &cur, 1); free(x); x = t;
        batchnorm(x, (size_t)n * hh * ww, c2s[l], &cur);
        cin = c2s[l];
    }
    t = conv_relu(x, n, hh, ww, cin, k, 1, &cur, 0); free(x); x = t;                        /* unet.py:63 */
    if (cur.i != n_weights) { free(x); return -1; }
    for (size_t p = 0; p < px0; ++p) {
        float *o = x + p * k;
        if (act == 0) {
            for (int j = 0; j < k; ++j) o[j] = 1.0f / (1.0f + expf(-o[j]));
        } else {
            float m = o[0], s = 0.f;
            for (int j = 1; j < k; ++j) if (o[j] > m) m = o[j];
            for (int j = 0; j < k; ++j) { o[j] = expf(o[j] - m); s += o[j]; }
            for (int j = 0; j < k; ++j) o[j] /= s;
        }
    }
    memcpy(probs, x, sizeof(float) * px0 * k);
    free(x);
    return 0;
}

/* functions.py:3104-3120 on int64 masks [M][P]; sizes = {im_size, pred_size} */
void oracle_im_binary(const int64_t *masks, int m, int64_t p, uint8_t *label, uint8_t *im, int64_t *sizes) {
    sizes[0] = sizes[1] = 0;
    for (int64_t i = 0; i < p; ++i) {
        int64_t s = 0;
        for (int j = 0; j < m; ++j) s += masks[(int64_t)j * p + i];
        const int all = s == m, mixed = (s != 0) && (s != m);
        label[i] = all ? 255 : 0;
        im[i] = mixed ? 255 : 0;
        sizes[0] += mixed;
        sizes[1] += all;
    }
}

/* functions.py:3123-3137 */
void oracle_im_multiclass(const int64_t *masks, int m, int64_t p, uint8_t *label, uint8_t *im, int64_t *sizes) {
    sizes[0] = 0;
    for (int64_t i = 0; i < p; ++i) {
        int agree = 1;
        for (int j = 1; j < m; ++j) agree &= masks[(int64_t)j * p + i] == masks[i];
        label[i] = agree ? (uint8_t)masks[i] : 0;
        im[i] = agree ? 0 : 255;
        sizes[0] += !agree;
    }
}
