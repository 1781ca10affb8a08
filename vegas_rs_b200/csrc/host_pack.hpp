// host_pack.hpp -- host side of the Ising State transfer for big lattices: the reference's host layout (one int8 +1 / -1 per
// spin, src/state.rs:60-63) <-> a sign bitmap in natural site order (bit i & 31 of word i >> 5 = spin i is Up), 1/8 of the
// bytes over PCIe.  Worker threads convert chunk by chunk so that the copies of finished chunks overlap the conversion of
// the next ones.  Pure host code (AVX2 when the CPU has it, 64-bit SWAR otherwise); no CUDA calls in here.
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <functional>

namespace vg {

// words[i] bit b = (s[32 i + b] > 0)
void host_pack_signs(const int8_t* s, uint32_t* words, size_t n_words);
// s[32 i + b] = bit ? +1 : -1
void host_unpack_signs(const uint32_t* words, int8_t* s, size_t n_words);

unsigned host_pack_threads();   // VEGAS_HOST_THREADS, else min(16, hardware_concurrency / LOCAL_WORLD_SIZE)

// Runs `work(chunk, first_word, n_words)` for every chunk in order on `threads` workers (each worker takes an equal share of
// every chunk) and calls `done(chunk)` on the CALLING thread as soon as all workers have finished that chunk; before a
// worker touches chunk c it waits until ready(c) returns true (polled; nullptr = always ready).
void host_chunked(size_t total_words, size_t chunk_words, unsigned threads,
                  const std::function<void(size_t first_word, size_t n_words)>& work,
                  const std::function<void(size_t chunk, size_t first_word, size_t n_words)>& done,
                  const std::function<void(size_t chunk)>& wait_ready);

}  // namespace vg
