/*
 * oip_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see oip_oracle.h).  Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off matters: the reference's default build has no FMA.
 *
 * "ref" = /root/reference/OpticalImageProcessor/<file>:<line>.
 */
#define _GNU_SOURCE
#include "oip_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ===================================================================================== */
/* stage 1                                                                               */
/* ===================================================================================== */

/* ref CRC.h:806-834 with CRC_16_CCITTFALSE = {0x1021, 0xFFFF, 0x0000, false, false} (:1522-1526),
 * Finalize (:707-720) is the identity for these parameters. */
uint16_t oipo_crc16(const uint8_t *data, size_t n)
{
    uint16_t rem = 0xFFFF;
    while (n--) {
        rem = (uint16_t)(rem ^ ((uint16_t)(*data++) << 8));
        for (int i = 0; i < 8; ++i)
            rem = (uint16_t)((rem & 0x8000) ? ((rem << 1) ^ 0x1021) : (rem << 1));
    }
    return rem;
}

static inline uint32_t be32(const uint8_t *p)
{
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
static inline uint32_t be16(const uint8_t *p) { return ((uint32_t)p[0] << 8) | p[1]; }

/* ref aux_separator.h:658-690 */
int oipo_aos_validate(const uint8_t *f, uint32_t *vcid_o, uint32_t *seq_o, uint32_t *inj_o,
                      uint32_t *crc_o)
{
    uint32_t vcid = f[5] & 0x3F;                                       /* :659 */
    uint32_t seq = ((uint32_t)f[6] << 16) | ((uint32_t)f[7] << 8) | f[8]; /* :660-661 */
    uint32_t inj = be32(f + 10);                                       /* :662-663 */
    uint32_t crc = be16(f + 894);                                      /* :665-666 */
    if (vcid_o) *vcid_o = vcid;
    if (seq_o) *seq_o = seq;
    if (inj_o) *inj_o = inj;
    if (crc_o) *crc_o = crc;
    if (inj != 0xAAAAAAAAu && inj != 0u) return -1;                    /* :675 */
    if (inj == 0xAAAAAAAAu && vcid == 0x3F) return 0;                  /* :676 */
    if (oipo_crc16(f + 4, 6 + 4 + 880) != crc) return -1;              /* :679-686 */
    return 1;
}

/* ref aux_separator.h:395-467 (loop :421-461), NextAosFrame :622-625 */
int64_t oipo_aos_scan(const uint8_t *buf, size_t n, uint64_t *payload_off, size_t cap,
                      int64_t counters[3])
{
    static const uint8_t sync[4] = {0x1A, 0xCF, 0xFC, 0x1D};
    int64_t valid = 0, invalid = 0, empty = 0;
    size_t p = 0, remain = n;
    for (;;) {
        if (remain < 1024) break;                                      /* :623 */
        const uint8_t *hit = (const uint8_t *)memmem(buf + p, remain, sync, 4); /* :624 */
        if (!hit) break;
        size_t h = (size_t)(hit - buf);
        if (h + 1024 > n) break; /* SURVEY C-5: the reference would read past the mapping */
        int r = oipo_aos_validate(hit, NULL, NULL, NULL, NULL);
        if (r != 1) {
            if (r < 0) invalid++; else empty++;                        /* :438-439 */
            remain -= (h - p) + 4;                                     /* :440-441 */
            p = h + 4;
            continue;
        }
        if ((size_t)valid < cap && payload_off) payload_off[valid] = h + 14; /* :458-460 */
        valid++;
        remain -= (h - p) + 1024;                                      /* :456-457 */
        p = h + 1024;
    }
    if (counters) { counters[0] = valid; counters[1] = invalid; counters[2] = empty; }
    return valid;
}

/* The same scan restricted to the candidates that start in [start, own_end) of a buffer of n bytes (a byte-range shard,
 * SURVEY 8e): the sequential loop of aux_separator.h:421-461 entered at search position `start` and left when the next
 * sync word starts at or after own_end.  *next_pos = where the scan would go on (>= own_end when a frame accepted here
 * reaches past the shard).  Running the shards one after the other, each from the previous one's next_pos, IS the
 * whole-file scan. */
int64_t oipo_aos_scan_range(const uint8_t *buf, size_t n, size_t start, size_t own_end, uint64_t *payload_off, size_t cap,
                            int64_t counters[3], uint64_t *next_pos)
{
    static const uint8_t sync[4] = {0x1A, 0xCF, 0xFC, 0x1D};
    int64_t valid = 0, invalid = 0, empty = 0;
    size_t p = start;
    for (;;) {
        if (p >= n || n - p < 1024) break;                             /* :623 */
        const uint8_t *hit = (const uint8_t *)memmem(buf + p, n - p, sync, 4);
        if (!hit) break;
        size_t h = (size_t)(hit - buf);
        if (h >= own_end) break;                                       /* the next shard's candidate */
        if (h + 1024 > n) break;
        int r = oipo_aos_validate(hit, NULL, NULL, NULL, NULL);
        if (r != 1) {
            if (r < 0) invalid++; else empty++;
            p = h + 4;
            continue;
        }
        if ((size_t)valid < cap && payload_off) payload_off[valid] = h + 14;
        valid++;
        p = h + 1024;
    }
    if (counters) { counters[0] = valid; counters[1] = invalid; counters[2] = empty; }
    if (next_pos) *next_pos = p;
    return valid;
}

/* ref aux_separator.h:469-556 (cadence :499-510), ValidateImtrFrame :558-590 */
int64_t oipo_imtr_deframe(const uint8_t *buf, const uint64_t *payload_off, int64_t n_payload,
                          uint8_t *imdt, size_t cap, int64_t stats[9])
{
    static const uint8_t sig[4] = {0x49, 0x54, 0xCE, 0x1F};
    static const uint8_t esig[4] = {0x2E, 0xE9, 0xC8, 0xFD};
    uint8_t cache[882 * 2];
    uint8_t frame[882];
    int cache_bytes = 0;
    int64_t next_payload = 0;
    uint32_t last_seq = 0;
    int64_t out = 0;
    int64_t st[9] = {0, 0, 0, 0, 0, 0, 0, -1, 0};
    for (;;) {
        if (cache_bytes < 882) {                                       /* :487 */
            if (next_payload >= n_payload) break;                      /* :495-498 sentinel */
            memcpy(cache + cache_bytes, buf + payload_off[next_payload++], 880); /* :499 */
            cache_bytes += 880;
            continue;
        }
        memcpy(frame, cache, 882);                                     /* :505 */
        cache_bytes -= 882;
        if (cache_bytes > 0) memmove(cache, cache + 882, (size_t)cache_bytes); /* :508 */
        st[0]++;
        /* ValidateImtrFrame */
        if (memcmp(frame, sig, 4) != 0) { st[2]++; continue; }         /* :559 */
        if (memcmp(frame + 878, esig, 4) != 0) { st[3]++; continue; }  /* :563 */
        uint32_t seq = be32(frame + 4);                                /* :568-569 */
        uint8_t chid = frame[8];
        if (frame[9] != 0x22) { st[4]++; continue; }                   /* :572 */
        if (oipo_crc16(frame, 876) != be16(frame + 876)) { st[5]++; continue; } /* :577-583 */
        st[1]++;
        if (last_seq == 0) {                                           /* :513-528: fopen "wb" */
            out = 0;
            st[7] = chid;
            st[8]++;
        }
        if (last_seq + 1 != seq) st[6]++;                              /* :530-533 */
        last_seq = seq;
        if ((size_t)out + 866 > cap) return -1;
        memcpy(imdt + out, frame + 10, 866);                           /* :536 */
        out += 866;
    }
    if (stats) memcpy(stats, st, sizeof st);
    return out;
}

/* ref aux_separator.h:256-393, :627-656 */
int64_t oipo_image_frames(const uint8_t *imdt, size_t n, const oipo_frame_geom *g, uint8_t *aux,
                          uint16_t *pan, uint16_t *mss, int64_t cap_frames, int64_t stats[4])
{
    static const uint8_t sig[4] = {0xEB, 0x90, 0xE1, 0x4D};
    const int TC = g->tile_cols, TL = g->tile_lines;
    const int64_t line_px = 8 * (int64_t)TC;
    const int64_t aux_all = 48 * 4 * (int64_t)TL;      /* IMGSIG_AUX_ALLBYTES */
    const int64_t tile_bytes = (int64_t)TL * TC * 2;   /* subImageBytes :342 */
    const int64_t pan_px = 4 * TL * line_px, mss_px = TL * line_px;
    size_t p = 0, remain = n;
    int last_seq = 0;
    int64_t emitted = 0, found = 0, incomplete = 0;
    for (;;) {
        /* NextImageDataFrame :627-656 */
        if (remain <= (size_t)(aux_all + 172)) break;                  /* :630 */
        const uint8_t *sp = (const uint8_t *)memmem(imdt + p, remain, sig, 4); /* :631 */
        if (!sp) break;
        size_t s = (size_t)(sp - imdt);
        if (s + 172 > n) break; /* the reference would parse past the mapping */
        size_t frame_end = s + 172;                                    /* :634 */
        uint8_t z_ratio = sp[4] & 0x3F;                                /* :639 */
        int seq = (int)be16(sp + 6);                                   /* :642-643 */
        uint32_t image_dwords = be32(sp + 8);                          /* :645-646 */
        uint32_t sub[40];
        for (int i = 0; i < 40; ++i) sub[i] = be32(sp + 12 + 4 * i);   /* :648-651 */
        int data_bytes = (int)(uint32_t)((uint64_t)image_dwords * 4u + (uint64_t)aux_all); /* :653 */
        if ((int64_t)(s - p) < (int64_t)data_bytes) {                  /* :654 -> :289-299 */
            incomplete++;
            remain -= frame_end - p;
            p = frame_end;
            continue;
        }
        if (data_bytes < 0) return -3;
        size_t frame = s - (size_t)data_bytes;                         /* :655 */
        found++;
        if (z_ratio != 0) return -2;
        if (seq > last_seq + 1) {                                      /* :302-311 */
            for (int i = 0; i < seq - last_seq - 1; ++i) {
                if (emitted < cap_frames) {
                    if (aux) memset(aux + emitted * aux_all, 0, (size_t)aux_all);
                    if (pan) memset(pan + emitted * pan_px, 0, (size_t)pan_px * 2);
                    if (mss) memset(mss + emitted * mss_px, 0, (size_t)mss_px * 2);
                }
                emitted++;
            }
        }
        if (emitted < cap_frames) {
            if (aux) memcpy(aux + emitted * aux_all, imdt + frame, (size_t)aux_all); /* :335-339 */
            /* WriteImageData :341-364 */
            size_t q = frame + (size_t)aux_all;
            for (int r = 0; r < 5; ++r) {
                for (int c = 0; c < 8; ++c) {
                    int idx = r * 8 + c;
                    if (q + (size_t)tile_bytes > n) return -3;
                    uint16_t *dst = (r < 4) ? (pan ? pan + emitted * pan_px + (int64_t)r * TL * line_px : NULL)
                                            : (mss ? mss + emitted * mss_px : NULL);
                    if (dst) {
                        for (int y = 0; y < TL; ++y) {                 /* MergeSubImage :366-372 */
                            const uint8_t *sl = imdt + q + (size_t)y * TC * 2;
                            uint16_t *dl = dst + (int64_t)y * line_px + (int64_t)c * TC;
                            for (int x = 0; x < TC; ++x)               /* swap :387-392 */
                                dl[x] = (uint16_t)(((uint16_t)sl[2 * x] << 8) | sl[2 * x + 1]);
                        }
                    }
                    q += (size_t)sub[idx] * 4u;                        /* :350,355 */
                }
            }
        }
        emitted++;
        remain -= frame_end - p;                                       /* :315-317 */
        p = frame_end;
        last_seq = seq;
    }
    if (stats) { stats[0] = found; stats[1] = emitted; stats[2] = incomplete; stats[3] = last_seq; }
    return emitted;
}

/* ===================================================================================== */
/* stage 2                                                                               */
/* ===================================================================================== */

/* ref imageop.h:129-138.  (uint16_t)(double) compiles to cvttsd2si + low 16 bits. */
void oipo_rrc_u16(uint16_t *buf, int w, int64_t h, const double *kb)
{
    for (int64_t y = 0; y < h; ++y) {
        uint16_t *row = buf + y * (int64_t)w;
        for (int x = 0; x < w; ++x) {
            uint16_t src = row[x];
            /* int conversion first: out-of-range double->uint16_t is UB in C; the reference's
             * x86 build goes through a 32-bit cvttsd2si, restated explicitly here */
            double v = kb[2 * x] * src + kb[2 * x + 1];
            int32_t t;
            if (v > -2147483649.0 && v < 2147483648.0) t = (int32_t)v;
            else t = (int32_t)0x80000000u; /* cvttsd2si "integer indefinite" */
            row[x] = (uint16_t)(t & 0xFFFF);
        }
    }
}

/* ref imageop.h:140-192 */
int oipo_load_rrc_csv(const char *path, int expected, double *kb)
{
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    char buff[1024];
    if (!fgets(buff, sizeof buff, f)) { fclose(f); return -2; }        /* :149 */
    if (!fgets(buff, sizeof buff, f)) { fclose(f); return -2; }        /* :156 */
    if (atoi(buff) != expected) { fclose(f); return -3; }              /* :159-162 */
    if (!fgets(buff, sizeof buff, f)) { fclose(f); return -2; }        /* :165 */
    int index = 0;
    double k, b;
    for (; fgets(buff, sizeof buff, f); ++index) {                     /* :177-183 */
        if (sscanf(buff, " %lf , %lf", &k, &b) != 2) { fclose(f); return -4; }
        if (index < expected) { kb[2 * index] = k; kb[2 * index + 1] = b; }
    }
    fclose(f);
    return index == expected ? 0 : -5;                                 /* :185-188 */
}

/* ref preproc.h:56-80 */
void oipo_mss_split(const uint16_t *mixed, int64_t lines, int line_px, uint16_t *planes[4])
{
    int bw = line_px / 4;
    for (int64_t i = 0; i < lines; ++i)
        for (int b = 0; b < 4; ++b)
            memcpy(planes[b] + i * bw, mixed + i * (int64_t)line_px + (int64_t)b * bw, (size_t)bw * 2);
}

/* ===================================================================================== */
/* stage 3: OpenCV cubic remap semantics                                                 */
/* ===================================================================================== */

/* OpenCV imgproc interpolateCubic (A=-0.75), float arithmetic, no contraction */
static void interpolate_cubic(float x, float *c)
{
    const float A = -0.75f;
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
}

void oipo_cubic_tab(float tab[32 * 4])
{
    const float scale = 1.f / 32;
    for (int i = 0; i < 32; ++i) interpolate_cubic(i * scale, tab + 4 * i);
}

static float g_tab1[32 * 4];
static int g_tab_ready = 0;
static void ensure_tab(void)
{
    if (!g_tab_ready) { oipo_cubic_tab(g_tab1); g_tab_ready = 1; }
}

static inline int cv_round(float v)
{
    /* cvRound: cvtss2si, round-half-even; out of range / NaN -> INT_MIN */
    if (!(v > -2147483904.0f && v < 2147483648.0f)) return (int)0x80000000u;
    return (int)lrintf(v);
}
static inline int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }
static inline uint16_t sat_u16_from_f(float v)
{
    int i = cv_round(v);
    return (uint16_t)(i < 0 ? 0 : (i > 65535 ? 65535 : i));
}

/* one output pixel of remapBicubic<Cast<float,ushort>,float,1>, BORDER_CONSTANT 0 */
static inline uint16_t remap_px(const uint16_t *src, int sw, int sh, int64_t sstep, float mx, float my)
{
    int sxf = cv_round(mx * 32.0f), syf = cv_round(my * 32.0f);
    int fx = sxf & 31, fy = syf & 31;
    int sx = sat_short(sxf >> 5) - 1, sy = sat_short(syf >> 5) - 1;
    const float *wx = g_tab1 + 4 * fx, *wy = g_tab1 + 4 * fy;
    unsigned width1 = (unsigned)(sw - 3 > 0 ? sw - 3 : 0), height1 = (unsigned)(sh - 3 > 0 ? sh - 3 : 0);
    if ((unsigned)sx < width1 && (unsigned)sy < height1) {
        const uint16_t *S = src + (int64_t)sy * sstep + sx;
        float w0 = wy[0] * wx[0], w1 = wy[0] * wx[1], w2 = wy[0] * wx[2], w3 = wy[0] * wx[3];
        float sum = S[0] * w0 + S[1] * w1 + S[2] * w2 + S[3] * w3;
        S += sstep;
        w0 = wy[1] * wx[0]; w1 = wy[1] * wx[1]; w2 = wy[1] * wx[2]; w3 = wy[1] * wx[3];
        sum += S[0] * w0 + S[1] * w1 + S[2] * w2 + S[3] * w3;
        S += sstep;
        w0 = wy[2] * wx[0]; w1 = wy[2] * wx[1]; w2 = wy[2] * wx[2]; w3 = wy[2] * wx[3];
        sum += S[0] * w0 + S[1] * w1 + S[2] * w2 + S[3] * w3;
        S += sstep;
        w0 = wy[3] * wx[0]; w1 = wy[3] * wx[1]; w2 = wy[3] * wx[2]; w3 = wy[3] * wx[3];
        sum += S[0] * w0 + S[1] * w1 + S[2] * w2 + S[3] * w3;
        return sat_u16_from_f(sum);
    }
    if (sx >= sw || sx + 4 <= 0 || sy >= sh || sy + 4 <= 0) return 0;
    float sum = 0.f;
    for (int i = 0; i < 4; ++i) {
        int yi = sy + i;
        if (yi < 0 || yi >= sh) continue;
        const uint16_t *S = src + (int64_t)yi * sstep;
        for (int k = 0; k < 4; ++k) {
            int xi = sx + k;
            if (xi >= 0 && xi < sw) sum += S[xi] * (wy[i] * wx[k]);
        }
    }
    return sat_u16_from_f(sum);
}

void oipo_remap_cubic_u16(const uint16_t *src, int sw, int sh, int64_t sstep, uint16_t *dst, int dw,
                          int dh, const float *mapx, const float *mapy)
{
    ensure_tab();
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            int64_t i = (int64_t)y * dw + x;
            dst[i] = remap_px(src, sw, sh, sstep, mapx[i], mapy[i]);
        }
}

/* dst local rows [j0,j1) of cv::remap(buff(hbuf x w), mapx = x+dX, mapy = y+dY). ref stitcher.h:93-99 */
static void shift_rows(const uint16_t *buff, int w, int hbuf, double dX, double dY, int j0, int j1,
                       uint16_t *out)
{
    ensure_tab();
    for (int y = j0; y < j1; ++y) {
        float my = (float)(y + dY);                                    /* :97 */
        uint16_t *o = out + (int64_t)(y - j0) * w;
        for (int x = 0; x < w; ++x) {
            float mx = (float)(x + dX);                                /* :96 */
            o[x] = remap_px(buff, w, hbuf, w, mx, my);
        }
    }
}

/* ref stitcher.h:83-139 + imageop.h:230-275 */
int64_t oipo_prestitch_shift(const uint16_t *src, int w, int64_t total_rows, double dX, double dY,
                             int section_rows, int row_guard, uint16_t *dst)
{
    int ucut = dY >= 0.0 ? 0 : (int)(-dY) + 1;                         /* stitcher.h:122 */
    int bcut = dY >= 0.0 ? (int)dY + 1 : 0;                            /* :123 */
    if (total_rows <= row_guard) {
        /* the reference throws here (imageop.h:242-244: "please use cv::remap()");
         * extension: exactly that -- one remap over the whole image, nothing cut */
        if (total_rows > 0) shift_rows(src, w, (int)total_rows, dX, dY, 0, (int)total_rows, dst);
        return total_rows;
    }
    if (section_rows > row_guard) return -1;
    int total_cut = ucut + bcut;                                       /* imageop.h:246 */
    if (total_cut >= section_rows) return -1;
    uint16_t *buff = (uint16_t *)calloc((size_t)section_rows * w, 2);  /* stitcher.h:88 */
    if (!buff) return -2;
    int64_t row_offset = 0, written = 0;
    int last_rows = 0;
    for (int s = 0;; ++s) {
        int64_t left = total_rows - row_offset;
        int rows = (int)(left < section_rows ? left : section_rows);   /* imageop.h:250 */
        if (rows <= total_cut) break;                                  /* :251 */
        memcpy(buff, src + row_offset * w, (size_t)rows * w * 2);      /* stitcher.h:105-106 */
        last_rows = rows;
        if (s == 0 && ucut > 0) {                                      /* imageop.h:260-263 */
            shift_rows(buff, w, section_rows, dX, dY, 0, ucut, dst + written * w);
            written += ucut;
        }
        shift_rows(buff, w, section_rows, dX, dY, ucut, rows - bcut, dst + written * w); /* :265 */
        written += rows - total_cut;
        row_offset += rows - total_cut;                                /* :266 */
    }
    (void)last_rows;
    if (bcut > 0) {                                                    /* :269-272 */
        shift_rows(buff, w, section_rows, dX, dY, section_rows - bcut, section_rows, dst + written * w);
        written += bcut;
    }
    free(buff);
    return written;
}

/* ref imageop.h:291-295, :340-355 generalised to n CCDs */
void oipo_stitch_concat_u16(const uint16_t *const *ccd, int n_ccd, int w, int64_t h, int f, uint16_t *dst)
{
    int64_t wout = (int64_t)n_ccd * w - 2 * (int64_t)(n_ccd - 1) * f;
    for (int64_t y = 0; y < h; ++y) {
        uint16_t *o = dst + y * wout;
        for (int i = 0; i < n_ccd; ++i) {
            int lo = i == 0 ? 0 : f, hi = i == n_ccd - 1 ? w : w - f;
            memcpy(o, ccd[i] + y * (int64_t)w + lo, (size_t)(hi - lo) * 2);
            o += hi - lo;
        }
    }
}

int64_t oipo_pan_pipeline(const uint16_t *const *ccd, int n_ccd, int w, int64_t h,
                          const double *const *kb, const double *dX, const double *dY, int f,
                          int section_rows, int row_guard, uint16_t *dst)
{
    uint16_t **tmp = (uint16_t **)calloc((size_t)n_ccd, sizeof *tmp);
    int64_t rc = h;
    for (int i = 0; i < n_ccd && rc >= 0; ++i) {
        uint16_t *r = (uint16_t *)malloc((size_t)h * w * 2);
        memcpy(r, ccd[i], (size_t)h * w * 2);
        if (kb && kb[i]) oipo_rrc_u16(r, w, h, kb[i]);                 /* Stitcher::DoRRC */
        if (i == 0) { tmp[i] = r; continue; }
        uint16_t *s = (uint16_t *)malloc((size_t)h * w * 2);
        int64_t wr = oipo_prestitch_shift(r, w, h, dX[i], dY[i], section_rows, row_guard, s);
        free(r);
        tmp[i] = s;
        if (wr != h) rc = -1;
    }
    if (rc >= 0) oipo_stitch_concat_u16((const uint16_t *const *)tmp, n_ccd, w, h, f, dst);
    for (int i = 0; i < n_ccd; ++i) free(tmp[i]);
    free(tmp);
    return rc;
}

/* ref preproc.h:428-468 : one section */
static void band_align_section(const uint16_t *const planes[4], int wb, int64_t row_offset, int rows,
                               const double cX[4][2], const double cY[4][3], uint16_t *out /* rows x wb x 4 */)
{
    ensure_tab();
    for (int b = 0; b < 4; ++b) {
        const double *coeffX = cX[b], *coeffY = cY[b];
        const uint16_t *src = planes[b] + row_offset * wb;            /* :453 */
        for (size_t y = 0; y < (size_t)rows; ++y) {
            for (int x = 0; x < wb; ++x) {
                size_t yy = y * 4;                                    /* :445 */
                int xx = x * 4;                                       /* :446 */
                float mx = (float)((coeffX[1] * xx + coeffX[0] + xx) / 4);                       /* :447 */
                float my = (float)((coeffY[2] * xx * xx + coeffY[1] * xx + coeffY[0] + yy) / 4); /* :448 */
                out[((int64_t)y * wb + x) * 4 + b] = remap_px(src, wb, rows, wb, mx, my);        /* :453-464 */
            }
        }
    }
}

static int g_min_process_lines = 1500; /* IBPA_MIN_PROCESSLINES oipshared.h:46 */
void oipo_set_min_process_lines(int v) { g_min_process_lines = v; }

/* ref preproc.h:351-425 */
int64_t oipo_band_align(const uint16_t *const planes[4], int64_t lines, int wb, const double cX[4][2],
                        const double cY[4][3], int lps, int64_t line_offset, int overlap, int keep,
                        uint16_t *out)
{
    if (overlap > 3000) return -1;                                     /* :355 */
    if (lps > 32767) return -2;                                        /* :359 */
    if (lps < overlap * 2) return -3;                                  /* :362 */
    if (lines - line_offset < g_min_process_lines) return -4;          /* :365 */
    uint64_t offset = (uint64_t)line_offset;
    int64_t processed = 0;
    uint16_t *sec = (uint16_t *)malloc((size_t)lps * wb * 4 * 2);
    if (!sec) return -5;
    for (int i = 0;; ++i) {
        uint64_t rem = (uint64_t)lines - offset;                       /* size_t arithmetic :380 */
        uint64_t n = rem < (uint64_t)lps ? rem : (uint64_t)lps;
        if ((uint64_t)lines < offset || n < (uint64_t)g_min_process_lines) break; /* :381 */
        band_align_section(planes, wb, (int64_t)offset, (int)n, cX, cY, sec);
        if (i == 0 && keep) {                                          /* :392-398 */
            memcpy(out, sec, (size_t)overlap * wb * 8);
            processed += overlap;
        }
        memcpy(out + processed * wb * 4, sec + (int64_t)overlap * wb * 4,
               (size_t)(n - (uint64_t)overlap) * wb * 8);              /* :400-402 */
        processed += (int64_t)n - overlap;                             /* :405 */
        offset += (uint64_t)(lps - overlap);                           /* :407 */
    }
    free(sec);
    return processed;
}

/* ref imageop.h:416-421 / :501-506 (geometry), :529 (band map) */
void oipo_stitch_concat_c4(const uint16_t *const *img, int n_img, int w, int64_t h, int f,
                           const int *band_map, uint16_t *dst)
{
    int64_t wout = (int64_t)n_img * w - 2 * (int64_t)(n_img - 1) * f;
    for (int64_t y = 0; y < h; ++y) {
        uint16_t *o = dst + y * wout * 4;
        for (int i = 0; i < n_img; ++i) {
            int lo = i == 0 ? 0 : f, hi = i == n_img - 1 ? w : w - f;
            const uint16_t *s = img[i] + (y * (int64_t)w + lo) * 4;
            for (int x = 0; x < hi - lo; ++x)
                for (int b = 0; b < 4; ++b)
                    o[x * 4 + b] = s[x * 4 + (band_map ? band_map[b] - 1 : b)];
            o += (int64_t)(hi - lo) * 4;
        }
    }
}

/* ===================================================================================== */
/* extension: packed samples                                                             */
/* ===================================================================================== */
void oipo_unpack_bits(const uint8_t *in, int bits, int w, int64_t h, int64_t pitch, uint16_t *out)
{
    for (int64_t y = 0; y < h; ++y) {
        const uint8_t *row = in + y * pitch;
        for (int x = 0; x < w; ++x) {
            int64_t bit = (int64_t)x * bits;
            uint32_t acc = ((uint32_t)row[bit >> 3] << 16);
            if (((bit + bits - 1) >> 3) >= (bit >> 3) + 1) acc |= ((uint32_t)row[(bit >> 3) + 1] << 8);
            if (((bit + bits - 1) >> 3) >= (bit >> 3) + 2) acc |= (uint32_t)row[(bit >> 3) + 2];
            out[y * (int64_t)w + x] = (uint16_t)((acc >> (24 - (bit & 7) - bits)) & ((1u << bits) - 1));
        }
    }
}

void oipo_swap16(const uint16_t *in, int64_t n, uint16_t *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = (uint16_t)((in[i] & 0x00FF) << 8 | (in[i] & 0xFF00) >> 8);
}
