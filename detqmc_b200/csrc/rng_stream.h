// Host-side random number stream of one replica.
//
// The reference drives every Metropolis decision of a replica from one RngWrapper = one
// dSFMT-19937 generator (rngwrapper.h:43-119).  The order in which values are consumed is part of
// the parity contract, and the number consumed per site is data dependent
// (rand01() is drawn only when the acceptance probability is <= 1, detsdwopdim.cpp:3113).
//
// RngStream keeps that stream on the host and lets the GPU consume it without a host round trip
// per site: the stream is a FIFO with unbounded look-ahead.  A sweep uploads a window of the next
// `w` values, the update kernels consume from it with a device-side cursor, and after the sweep the
// host advances the FIFO by the number the device reports.  Values that were pre-drawn but not
// consumed stay at the head of the FIFO, so the logical stream is exactly the reference's.
#pragma once

#include <cstddef>
#include <cstdint>
#include <deque>
#include <vector>

namespace dqmc {

// dSFMT-19937, double precision SIMD-oriented Fast Mersenne Twister (Saito & Matsumoto 2009).
// Written from the published recursion; bit-exact with the reference's vendored dSFMT 2.1
// (src/dsfmt/dSFMT.c) for dsfmt_init_gen_rand + dsfmt_genrand_open_open.
class Dsfmt19937 {
public:
    static constexpr int kWords128 = 191;              // (19937 - 128) / 104 + 1
    static constexpr int kDoubles = 2 * kWords128;     // 382 values per regeneration
    explicit Dsfmt19937(uint32_t seed = 0) { seed_with(seed); }
    void seed_with(uint32_t seed);
    // uniform in (0, 1)
    double next_open_open();
    // raw state access for checkpoints: kDoubles + 2 words and the read index
    const uint64_t* state() const { return st_; }
    uint64_t* state() { return st_; }
    int index() const { return idx_; }
    void set_index(int i) { idx_ = i; }
private:
    void regenerate();
    uint64_t st_[kDoubles + 2];
    int idx_;
};

// seed scrambling of RngWrapper's constructor (rngwrapper.cpp:43), uint32 wrap-around arithmetic
uint32_t scramble_seed(uint32_t seed, uint32_t process_index);

typedef void (*rng_fill_fn)(void* user, double* out, size_t n);

class RngStream {
public:
    RngStream() : gen_(0), fill_(nullptr), user_(nullptr), consumed_(0) {}
    void seed(uint32_t seed, uint32_t process_index);
    void set_source(rng_fill_fn fill, void* user);
    // make sure at least n values are buffered and return a pointer to the head
    const double* peek(size_t n);
    void skip(size_t n);
    double draw();                                     // rand01()
    double draw_range(double lo, double hi) { return lo + (hi - lo) * draw(); }
    int draw_int(int lo, int hi) { return lo + static_cast<int>((hi - lo + 1.0) * draw()); }
    uint64_t consumed() const { return consumed_; }
    // make sure at least n values are buffered (generation can then overlap with device work)
    void prefetch(size_t n) { ensure(n); }
    // Read-only window: between begin_window(off) and end_window() draws return the values at head + off + i WITHOUT
    // consuming them -- that stretch of the stream is resident on the device, whose cursor owns the consumption.
    // end_window() returns how many values were read.
    void begin_window(size_t off) { win_ = true; winOff_ = off; winUsed_ = 0; }
    size_t end_window() { win_ = false; return winUsed_; }
    // values drawn from the source ahead of consumption (part of the replica's state in a checkpoint)
    size_t buffered() const { return buf_.size() - head_; }
    const double* buffered_values() const { return buf_.data() + head_; }
    // resume: replace the buffered values (drawn from a superseded generator state) by the checkpointed look-ahead
    void set_buffered(const double* v, size_t n) { buf_.assign(v, v + n); head_ = 0; }
private:
    void ensure(size_t n);
    Dsfmt19937 gen_;
    rng_fill_fn fill_;
    void* user_;
    std::vector<double> buf_;     // buffered values [head_, buf_.size())
    size_t head_ = 0;
    uint64_t consumed_;
    bool win_ = false;
    size_t winOff_ = 0, winUsed_ = 0;
};

}  // namespace dqmc
