// Host check of csrc/crc_bitslice.cuh: the bit-sliced 32-frame CRC against the bit-serial definition
// (ref CRC.h:806-834, parameters :1522-1526), 32 emulated lanes.  Built and run by tests/test_crc_bitslice_cpu.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../opticalimageprocessor_b200/csrc/crc_bitslice.cuh"

using namespace oip::bitslice;

static uint16_t crc_serial(const uint8_t *p, int n)
{
    uint16_t r = 0xFFFF;
    for (int i = 0; i < n; ++i)
        for (int b = 7; b >= 0; --b) {
            const int in = (p[i] >> b) & 1;
            const int top = (r >> 15) & 1;
            r = (uint16_t)(r << 1);
            if (top ^ in) r ^= 0x1021;
        }
    return r;
}

static uint64_t rng = 0x9E3779B97F4A7C15ull;
static uint32_t rnd()
{
    rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
    return (uint32_t)(rng >> 16);
}

template <int LEN> static int run(int trials)
{
    constexpr int HEAD = SPAN - LEN;
    int bad = 0;
    for (int tr = 0; tr < trials; ++tr) {
        // 32 frames at arbitrary byte offsets inside one buffer, garbage everywhere else
        std::vector<uint8_t> buf(32 * 1100 + 64);
        for (auto &b : buf) b = (uint8_t)rnd();
        int start[32];
        for (int f = 0; f < 32; ++f) start[f] = 32 + f * 1100 + (tr == 0 ? 0 : (int)(rnd() % 7)); // message start
        if (tr == 1) for (int f = 0; f < 32; ++f) memset(&buf[start[f]], 0, LEN);
        uint32_t P[32][16];
        for (int lane = 0; lane < 32; ++lane) {
            auto load = [&](int j, uint32_t(&T)[32]) {
                for (int f = 0; f < 32; ++f) {
                    const uint8_t *q = &buf[start[f] - HEAD + PIECE * lane + 4 * j];
                    T[f] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
                }
            };
            piece32(load, lane, HEAD, P[lane]);
        }
        // butterfly with an explicit exchange buffer
        auto level = [&](auto tag) {
            constexpr int S = decltype(tag)::value;
            uint32_t sent[32][16];
            for (int lane = 0; lane < 32; ++lane) {
                uint32_t M[16];
                mul_xpow<8 * PIECE * S>(P[lane], M);
                for (int i = 0; i < 16; ++i) sent[lane][i] = (lane & S) == 0 ? M[i] : P[lane][i];
            }
            for (int lane = 0; lane < 32; ++lane)
                for (int i = 0; i < 16; ++i) P[lane][i] = sent[lane][i] ^ sent[lane ^ S][i];
        };
        level(std::integral_constant<int, 1>{});
        level(std::integral_constant<int, 2>{});
        level(std::integral_constant<int, 4>{});
        level(std::integral_constant<int, 8>{});
        level(std::integral_constant<int, 16>{});
        for (int f = 0; f < 32; ++f) {
            const uint16_t want = crc_serial(&buf[start[f]], LEN);
            for (int lane = 0; lane < 32; lane += 31) {
                const uint16_t got = (uint16_t)(unslice(P[lane], f) ^ init_term(LEN));
                if (got != want) {
                    if (bad < 5) printf("LEN %d trial %d frame %d lane %d: got %04X want %04X\n", LEN, tr, f, lane, got, want);
                    ++bad;
                }
            }
        }
    }
    return bad;
}

// word-interleaved pieces (the global-memory variant): lane l = word l of each of the seven 128-byte blocks
template <int S> static void join_all32(uint32_t (&P)[32][16])
{
    uint32_t sent[32][16];
    for (int lane = 0; lane < 32; ++lane) {
        uint32_t M[16];
        mul_xpow<32 * S>(P[lane], M);
        for (int i = 0; i < 16; ++i) sent[lane][i] = (lane & S) == 0 ? M[i] : P[lane][i];
    }
    for (int lane = 0; lane < 32; ++lane) {
        auto x = [&](uint32_t, int s, int i) { return sent[lane ^ s][i]; };
        join_level<S, 32>(P[lane], lane, x);
    }
}
template <int LEN> static int run_interleaved(int trials)
{
    constexpr int HEAD = SPAN - LEN;
    int bad = 0;
    for (int tr = 0; tr < trials; ++tr) {
        std::vector<uint8_t> buf(32 * 1100 + 64);
        for (auto &b : buf) b = (uint8_t)rnd();
        int start[32];
        for (int f = 0; f < 32; ++f) start[f] = 32 + f * 1100 + (tr == 0 ? 0 : (int)(rnd() % 7));
        uint32_t P[32][16];
        for (int lane = 0; lane < 32; ++lane) {
            auto load = [&](int j, uint32_t(&T)[32]) {
                for (int f = 0; f < 32; ++f) {
                    const uint8_t *q = &buf[start[f] - HEAD + 128 * j + 4 * lane];
                    T[f] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
                }
            };
            piece32_interleaved(load, lane, HEAD, P[lane]);
        }
        join_all32<1>(P); join_all32<2>(P); join_all32<4>(P); join_all32<8>(P); join_all32<16>(P);
        for (int f = 0; f < 32; ++f) {
            const uint16_t want = crc_serial(&buf[start[f]], LEN);
            const uint16_t got = (uint16_t)(unslice(P[5], f) ^ init_term(LEN));
            if (got != want) {
                if (bad < 5) printf("interleaved LEN %d trial %d frame %d: got %04X want %04X\n", LEN, tr, f, got, want);
                ++bad;
            }
        }
    }
    return bad;
}

// join_level itself (the device code path) with an emulated exchange: two passes per level
template <int S> static void join_all(uint32_t (&P)[32][16])
{
    uint32_t sent[32][16];
    for (int lane = 0; lane < 32; ++lane) {
        uint32_t M[16];
        mul_xpow<8 * PIECE * S>(P[lane], M);
        for (int i = 0; i < 16; ++i) sent[lane][i] = (lane & S) == 0 ? M[i] : P[lane][i];
    }
    for (int lane = 0; lane < 32; ++lane) {
        auto x = [&](uint32_t, int s, int i) { return sent[lane ^ s][i]; };
        join_level<S>(P[lane], lane, x);
    }
}

int main()
{
    if (crc_serial((const uint8_t *)"123456789", 9) != 0x29B1) { puts("reference CRC self-check failed"); return 2; } // CRC.h:1519
    // transpose orientation
    uint32_t a[32], b[32];
    for (int i = 0; i < 32; ++i) a[i] = b[i] = rnd();
    transpose32(a);
    for (int k = 0; k < 32; ++k)
        for (int f = 0; f < 32; ++f)
            if (((a[k] >> f) & 1) != ((b[f] >> k) & 1)) { puts("transpose32 wrong"); return 3; }
    int bad = run<890>(20) + run<876>(20) + run_interleaved<890>(20) + run_interleaved<876>(20) + run_interleaved<892>(20);
    { // message followed by its own CRC (big-endian) leaves a zero remainder: what aos_crc_kernel tests
        uint8_t m[892];
        for (int i = 0; i < 890; ++i) m[i] = (uint8_t)rnd();
        const uint16_t c = crc_serial(m, 890);
        m[890] = (uint8_t)(c >> 8); m[891] = (uint8_t)c;
        if (crc_serial(m, 892) != 0) { puts("msg||crc remainder not zero"); ++bad; }
        m[100] ^= 0x10;
        if (crc_serial(m, 892) == 0) { puts("corrupted msg||crc remainder zero"); ++bad; }
    }
    // join_level against the open-coded butterfly
    {
        uint32_t P[32][16], Q[32][16];
        for (int l = 0; l < 32; ++l) for (int i = 0; i < 16; ++i) P[l][i] = Q[l][i] = rnd();
        join_all<4>(P);
        for (int lane = 0; lane < 32; ++lane) {
            uint32_t M[16], N[16];
            mul_xpow<8 * PIECE * 4>(Q[lane], M);
            mul_xpow<8 * PIECE * 4>(Q[lane ^ 4], N);
            for (int i = 0; i < 16; ++i) {
                const uint32_t want = (lane & 4) == 0 ? (M[i] ^ Q[lane ^ 4][i]) : (Q[lane][i] ^ N[i]);
                if (P[lane][i] != want) ++bad;
            }
        }
    }
    printf("mismatches: %d\n", bad);
    return bad ? 1 : 0;
}
