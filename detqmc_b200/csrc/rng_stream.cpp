#include "rng_stream.h"

#include <cstring>

namespace dqmc {

namespace {
constexpr int kPos1 = 117;
constexpr int kShiftLeft = 19;
constexpr int kShiftRight = 12;
constexpr uint64_t kMask1 = 0x000ffafffffffb3fULL;
constexpr uint64_t kMask2 = 0x000ffdfffc90fffdULL;
constexpr uint64_t kFix1 = 0x90014964b32f4329ULL;
constexpr uint64_t kFix2 = 0x3b8d12ac548a7c7aULL;
constexpr uint64_t kPcv1 = 0x3d84e1ac0dc82880ULL;
constexpr uint64_t kPcv2 = 0x0000000000000001ULL;
constexpr uint64_t kLowMask = 0x000FFFFFFFFFFFFFULL;
constexpr uint64_t kHighConst = 0x3FF0000000000000ULL;
}  // namespace

void Dsfmt19937::seed_with(uint32_t seed) {
    // linear congruential fill of the state viewed as 32-bit words (little endian pairing)
    constexpr int n32 = (kWords128 + 1) * 4;
    uint32_t w[n32];
    w[0] = seed;
    for (int i = 1; i < n32; ++i) {
        uint32_t prev = w[i - 1];
        w[i] = 1812433253u * (prev ^ (prev >> 30)) + static_cast<uint32_t>(i);
    }
    for (int i = 0; i < n32 / 2; ++i) {
        st_[i] = static_cast<uint64_t>(w[2 * i]) | (static_cast<uint64_t>(w[2 * i + 1]) << 32);
    }
    // force the IEEE exponent of [1, 2) on the value words
    for (int i = 0; i < kDoubles; ++i) st_[i] = (st_[i] & kLowMask) | kHighConst;
    // period certification on the two trailing "lung" words
    uint64_t inner = ((st_[kDoubles] ^ kFix1) & kPcv1) ^ ((st_[kDoubles + 1] ^ kFix2) & kPcv2);
    for (int sh = 32; sh > 0; sh >>= 1) inner ^= inner >> sh;
    if ((inner & 1) == 0) st_[kDoubles + 1] ^= 1;       // lowest bit of kPcv2 is set
    idx_ = kDoubles;
}

void Dsfmt19937::regenerate() {
    uint64_t l0 = st_[kDoubles], l1 = st_[kDoubles + 1];
    for (int i = 0; i < kWords128; ++i) {
        int j = i + kPos1;
        if (j >= kWords128) j -= kWords128;
        const uint64_t a0 = st_[2 * i], a1 = st_[2 * i + 1];
        const uint64_t n0 = (a0 << kShiftLeft) ^ (l1 >> 32) ^ (l1 << 32) ^ st_[2 * j];
        const uint64_t n1 = (a1 << kShiftLeft) ^ (l0 >> 32) ^ (l0 << 32) ^ st_[2 * j + 1];
        l0 = n0;
        l1 = n1;
        st_[2 * i] = (l0 >> kShiftRight) ^ (l0 & kMask1) ^ a0;
        st_[2 * i + 1] = (l1 >> kShiftRight) ^ (l1 & kMask2) ^ a1;
    }
    st_[kDoubles] = l0;
    st_[kDoubles + 1] = l1;
}

double Dsfmt19937::next_open_open() {
    if (idx_ >= kDoubles) {
        regenerate();
        idx_ = 0;
    }
    uint64_t bits = st_[idx_++] | 1;
    double d;
    std::memcpy(&d, &bits, sizeof d);
    return d - 1.0;
}

uint32_t scramble_seed(uint32_t seed, uint32_t process_index) {
    return ((seed * 181u) * ((process_index - 83u) * 359u)) % 104729u;
}

void RngStream::seed(uint32_t seed, uint32_t process_index) {
    gen_.seed_with(scramble_seed(seed, process_index));
    fill_ = nullptr;
    user_ = nullptr;
    buf_.clear();
    head_ = 0;
    consumed_ = 0;
}

void RngStream::set_source(rng_fill_fn fill, void* user) {
    fill_ = fill;
    user_ = user;
    buf_.clear();
    head_ = 0;
}

void RngStream::ensure(size_t n) {
    size_t have = buf_.size() - head_;
    if (have >= n) return;
    if (head_ > 0 && head_ >= buf_.size() / 2) {        // compact
        buf_.erase(buf_.begin(), buf_.begin() + static_cast<std::ptrdiff_t>(head_));
        head_ = 0;
    }
    size_t need = n - have;
    size_t old = buf_.size();
    buf_.resize(old + need);
    if (fill_) {
        fill_(user_, buf_.data() + old, need);
    } else {
        for (size_t i = 0; i < need; ++i) buf_[old + i] = gen_.next_open_open();
    }
}

const double* RngStream::peek(size_t n) {
    ensure(n);
    return buf_.data() + head_;
}

void RngStream::skip(size_t n) {
    ensure(n);
    head_ += n;
    consumed_ += n;
}

double RngStream::draw() {
    if (win_) {
        ensure(winOff_ + winUsed_ + 1);
        return buf_[head_ + winOff_ + winUsed_++];
    }
    ensure(1);
    consumed_ += 1;
    return buf_[head_++];
}

}  // namespace dqmc
