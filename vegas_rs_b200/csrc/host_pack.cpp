// host_pack.cpp -- see host_pack.hpp.  Replaces nothing in the reference: it is the marshalling of State<IsingSpin>
// (src/state.rs:245-318, one byte per spin) at the C-ABI boundary (vegas_gpu_upload_ising / _download_ising /
// _step_host_ising) for lattices where PCIe time dominates.
#include "host_pack.hpp"

#include <immintrin.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace vg {

namespace {

__attribute__((target("avx2"))) void pack_avx2(const int8_t* s, uint32_t* words, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 32 * i));
        // Up = +1: sign bit clear and non-zero; the reference's IsingSpin has only +1 / -1, anything <= 0 counts as Down
        const __m256i up = _mm256_cmpgt_epi8(v, _mm256_setzero_si256());
        words[i] = (uint32_t)_mm256_movemask_epi8(up);
    }
}

__attribute__((target("avx2"))) void unpack_avx2(const uint32_t* words, int8_t* s, size_t n) {
    const __m256i sel = _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3);
    const __m256i bitm = _mm256_set1_epi64x((long long)0x8040201008040201ull);
    const __m256i one = _mm256_set1_epi8(1), minus = _mm256_set1_epi8(-1);
    const bool aligned = (reinterpret_cast<uintptr_t>(s) & 31u) == 0;
    for (size_t i = 0; i < n; ++i) {
        const __m256i b = _mm256_shuffle_epi8(_mm256_set1_epi32((int)words[i]), sel);
        const __m256i up = _mm256_cmpeq_epi8(_mm256_and_si256(b, bitm), bitm);
        const __m256i v = _mm256_blendv_epi8(minus, one, up);
        if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i*>(s + 32 * i), v);   // the State is written once, never re-read here
        else _mm256_storeu_si256(reinterpret_cast<__m256i*>(s + 32 * i), v);
    }
    if (aligned) _mm_sfence();
}

void pack_swar(const int8_t* s, uint32_t* words, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        uint32_t w = 0;
        for (int k = 0; k < 4; ++k) {
            uint64_t v;
            memcpy(&v, s + 32 * i + 8 * k, 8);
            // byte > 0  <=>  sign bit clear and byte != 0; gather the eight flags with one multiply
            const uint64_t nz = ((v & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | v;   // bit 7 of a byte set iff byte != 0
            const uint64_t up = nz & ~v & 0x8080808080808080ull;
            w |= (uint32_t)(((up >> 7) * 0x0102040810204080ull) >> 56) << (8 * k);
        }
        words[i] = w;
    }
}

void unpack_swar(const uint32_t* words, int8_t* s, size_t n) {
    for (size_t i = 0; i < n; ++i)
        for (int b = 0; b < 32; ++b) s[32 * i + b] = (words[i] >> b) & 1u ? 1 : -1;
}

bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}

}  // namespace

void host_pack_signs(const int8_t* s, uint32_t* words, size_t n_words) {
    if (have_avx2()) pack_avx2(s, words, n_words); else pack_swar(s, words, n_words);
}
void host_unpack_signs(const uint32_t* words, int8_t* s, size_t n_words) {
    if (have_avx2()) unpack_avx2(words, s, n_words); else unpack_swar(words, s, n_words);
}

unsigned host_pack_threads() {
    if (const char* e = std::getenv("VEGAS_HOST_THREADS")) {
        const long v = std::strtol(e, nullptr, 10);
        if (v >= 1 && v <= 256) return (unsigned)v;
    }
    unsigned hc = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) {   // one process per GPU (torchrun): the ranks share the host's cores
        const long r = std::strtol(e, nullptr, 10);
        if (r > 1 && r <= 64) hc = std::max(1u, hc / (unsigned)r);
    }
    return std::max(1u, std::min(16u, hc ? hc : 1u));
}

void host_chunked(size_t total_words, size_t chunk_words, unsigned threads,
                  const std::function<void(size_t, size_t)>& work,
                  const std::function<void(size_t, size_t, size_t)>& done,
                  const std::function<void(size_t)>& wait_ready) {
    chunk_words = std::max<size_t>(1, chunk_words);
    const size_t n_chunks = (total_words + chunk_words - 1) / chunk_words;
    threads = std::max(1u, threads);
    if (threads == 1) {
        for (size_t c = 0; c < n_chunks; ++c) {
            const size_t first = c * chunk_words, nw = std::min(chunk_words, total_words - first);
            if (wait_ready) wait_ready(c);
            work(first, nw);
            if (done) done(c, first, nw);
        }
        return;
    }
    // a lattice of a few chunks' worth of words does not pay for thread start-up
    if (total_words < 4096) threads = 1;
    if (threads == 1) { host_chunked(total_words, chunk_words, 1, work, done, wait_ready); return; }
    std::vector<std::atomic<unsigned>> finished(n_chunks);
    std::vector<std::atomic<unsigned>> ready(n_chunks);
    for (size_t c = 0; c < n_chunks; ++c) { finished[c].store(0); ready[c].store(wait_ready ? 0u : 1u); }
    std::vector<std::thread> pool;
    pool.reserve(threads);
    std::atomic<int> go(0);   // 1: all workers exist, start; 2: a worker could not be created, leave
    auto single = [&]() {
        for (size_t c = 0; c < n_chunks; ++c) {
            const size_t first = c * chunk_words, nw = std::min(chunk_words, total_words - first);
            if (wait_ready) wait_ready(c);
            work(first, nw);
            if (done) done(c, first, nw);
        }
    };
    try {
    for (unsigned t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            int g;
            while ((g = go.load(std::memory_order_acquire)) == 0) std::this_thread::yield();
            if (g == 2) return;
            for (size_t c = 0; c < n_chunks; ++c) {
                while (ready[c].load(std::memory_order_acquire) == 0u) std::this_thread::yield();
                const size_t first = c * chunk_words, nw = std::min(chunk_words, total_words - first);
                const size_t per = (nw + threads - 1) / threads, a = std::min(nw, (size_t)t * per), b = std::min(nw, a + per);
                if (b > a) work(first + a, b - a);
                finished[c].fetch_add(1u, std::memory_order_release);
            }
        });
    } catch (...) {   // out of threads: the workers that exist leave, the caller does the work itself
        go.store(2, std::memory_order_release);
        for (auto& th : pool) th.join();
        single();
        return;
    }
    go.store(1, std::memory_order_release);
    // the calling thread owns the CUDA side: it opens chunks (wait_ready) and hands finished ones on (done), in order
    size_t opened = 0;
    for (size_t c = 0; c < n_chunks; ++c) {
        // keep one chunk open ahead of the one being waited for, so that the workers never idle on the caller
        for (; opened < n_chunks && opened <= c + 1; ++opened) {
            if (wait_ready) { wait_ready(opened); ready[opened].store(1u, std::memory_order_release); }
        }
        while (finished[c].load(std::memory_order_acquire) < threads) std::this_thread::yield();
        if (done) done(c, c * chunk_words, std::min(chunk_words, total_words - c * chunk_words));
    }
    for (auto& th : pool) th.join();
}

}  // namespace vg
