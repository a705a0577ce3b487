/* CPU oracle helpers in plain C -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The same pure-integer functions as oracle/synth.py (synth_frames) and oracle/degrade.py
 * (add_noise), restated in C only because the NumPy forms need minutes per 1080p / 1800-frame clip
 * and the configuration goldens (tests/golden/make_config_golden.py) regenerate 64 of them.  The
 * reference has no generator (its videos are git-ignored, /root/reference/.gitignore:1-5); the
 * noise degradation follows /root/reference/analysis/degradation/colour_noise.py:11-24 with the
 * np.random.normal draw replaced by the counter-based 12-term Irwin-Hall sum documented in
 * oracle/degrade.py.  tests/test_oracle.py holds both functions bit-exactly to the NumPy forms.
 *
 * Built by oracle/build.py:  gcc -O2 -shared -fPIC -o oracle/_build/liboracle_c.so oracle/csrc/synth_ref.c
 */
#include <stdint.h>

static inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

/* frames t0 .. t0+n-1 of a clip; pulse_q8 (T_total,3) int32; base_q8 [2][3]; face x0,y0,x1,y1 half-open */
void oracle_synth_frames(uint32_t seed, uint32_t clip, int t0, int n, int H, int W, const int32_t* face,
                         const int32_t* base_q8, int32_t noise_gain, const int32_t* pulse_q8, uint8_t* out) {
    const int64_t frame_bytes = (int64_t)H * W * 3;
    for (int i = 0; i < n; ++i) {
        const uint32_t t = (uint32_t)(t0 + i);
        const uint32_t key = mix32(seed * 0x9E3779B1u + clip * 0x85EBCA77u + t * 0xC2B2AE3Du + 0x165667B1u);
        const int32_t* pq = pulse_q8 + (int64_t)(t0 + i) * 3;
        uint8_t* o = out + (int64_t)i * frame_bytes;
        uint32_t idx = 0;
        for (int y = 0; y < H; ++y) {
            for (int x = 0; x < W; ++x) {
                const int f = (x >= face[0]) & (x < face[2]) & (y >= face[1]) & (y < face[3]);
                for (int c = 0; c < 3; ++c, ++idx) {
                    const uint32_t r = mix32(key ^ (idx * 0x27D4EB2Fu));
                    const int s = (int)((r & 255u) + ((r >> 8) & 255u) + ((r >> 16) & 255u) + (r >> 24)) - 510;
                    int v = base_q8[f * 3 + c] + f * pq[c] + s * noise_gain;
                    v = (v + 128) >> 8;
                    o[idx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
                }
            }
        }
    }
}

/* in place allowed; gain_q16 per unit of the centred twelve-byte sum */
void oracle_add_noise(const uint8_t* in, uint8_t* out, int T, int64_t frame_bytes, int32_t gain_q16, uint32_t seed,
                      uint32_t clip, int t0) {
    for (int i = 0; i < T; ++i) {
        const uint32_t key = mix32(seed * 0x9E3779B1u + clip * 0x85EBCA77u + (uint32_t)(t0 + i) * 0xC2B2AE3Du + 0x3C6EF372u);
        const uint8_t* p = in + (int64_t)i * frame_bytes;
        uint8_t* o = out + (int64_t)i * frame_bytes;
        for (uint32_t idx = 0; idx < (uint32_t)frame_bytes; ++idx) {
            int s = -1530;
            for (uint32_t w = 0; w < 3; ++w) {
                const uint32_t r = mix32((key + w * 0x9E3779B9u) ^ (idx * 0x27D4EB2Fu));
                s += (int)((r & 255u) + ((r >> 8) & 255u) + ((r >> 16) & 255u) + (r >> 24));
            }
            int v = ((int)p[idx] * 65536 + s * gain_q16) >> 16;
            o[idx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
}
