// Host-side probe: how fast can the box's cores turn the reference's int8 (+1 / -1) Ising State into a sign bitmap and back?
// (decides whether vegas_gpu_step_host_ising should pack on the host: 1/8 of the PCIe bytes, but the cores must stream 1 GiB)
// g++ -O3 -mavx2 -pthread host_pack_probe.cpp -o host_pack_probe && ./host_pack_probe
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void pack(const int8_t* s, uint32_t* bits, size_t n32) {   // bit = 1 for Up (+1): sign bit clear
    for (size_t i = 0; i < n32; ++i) {
        const __m256i v = _mm256_loadu_si256((const __m256i*)(s + 32 * i));
        bits[i] = ~(uint32_t)_mm256_movemask_epi8(v);
    }
}
static void unpack(const uint32_t* bits, int8_t* s, size_t n32) {
    const __m256i sel = _mm256_setr_epi8(0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,2,2,2,2,2,2,2,2,3,3,3,3,3,3,3,3);
    const __m256i bitm = _mm256_set1_epi64x((long long)0x8040201008040201ull);
    const __m256i one = _mm256_set1_epi8(1), minus = _mm256_set1_epi8(-1);
    for (size_t i = 0; i < n32; ++i) {
        const __m256i b = _mm256_shuffle_epi8(_mm256_set1_epi32((int)bits[i]), sel);
        const __m256i up = _mm256_cmpeq_epi8(_mm256_and_si256(b, bitm), bitm);
        _mm256_stream_si256((__m256i*)(s + 32 * i), _mm256_blendv_epi8(minus, one, up));
    }
}
int main() {
    const size_t n = (size_t)1 << 30;
    int8_t* s = (int8_t*)aligned_alloc(4096, n);
    uint32_t* bits = (uint32_t*)aligned_alloc(4096, n / 8);
    for (size_t i = 0; i < n; ++i) s[i] = (i * 2654435761u >> 13) & 1 ? 1 : -1;
    printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
    for (unsigned T : {1u, 2u, 4u, 8u, 16u, 32u, 64u}) {
        if (T > 2 * std::thread::hardware_concurrency()) break;
        for (int what = 0; what < 2; ++what) {
            double best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                const size_t n32 = n / 32, per = (n32 + T - 1) / T;
                for (unsigned t = 0; t < T; ++t) {
                    const size_t a = t * per, b = std::min(n32, a + per);
                    if (a >= b) break;
                    if (what == 0) th.emplace_back(pack, s + 32 * a, bits + a, b - a);
                    else th.emplace_back(unpack, bits + a, s + 32 * a, b - a);
                }
                for (auto& x : th) x.join();
                best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            }
            printf("%s threads %2u: %.2f ms for 1 GiB of spins = %.1f GB/s\n", what ? "unpack" : "pack  ", T, best * 1e3, n / best / 1e9);
        }
    }
    size_t chk = 0; for (size_t i = 0; i < n; i += 4097) chk += s[i] > 0; printf("check %zu\n", chk);
    return 0;
}
