// Recycling allocator for the per-image host buffers of the parser (codestream copy, AC code blob, coefficient orders):
// blocks of 64 KB and more come from a process-wide pool of power-of-two size classes instead of malloc, which serves
// such sizes with mmap / munmap -- 64 images per batch on several threads then spend more time in page faults and on
// the kernel's per-process mapping lock than in parsing.  Contents of a recycled block are unspecified, as with malloc.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace jxlb {

class BlockPool {
 public:
  static constexpr size_t kMinBytes = 64u << 10;
  static constexpr size_t kMaxRetained = (size_t) 2 << 30;  // beyond this, freed blocks go back to malloc
  static void* Get(size_t bytes, size_t* cls_bytes) {
    int c = 0;
    size_t cb = kMinBytes;
    while (cb < bytes) {
      cb <<= 1;
      ++c;
    }
    *cls_bytes = cb;
    if (c < kClasses) {
      std::lock_guard<std::mutex> l(Mu());
      auto& fl = Free()[c];
      if (!fl.empty()) {
        void* p = fl.back();
        fl.pop_back();
        Retained() -= cb;
        return p;
      }
    }
    void* p = malloc(cb);
    if (!p) throw std::bad_alloc();
    return p;
  }
  static void Put(void* p, size_t bytes) {
    int c = 0;
    size_t cb = kMinBytes;
    while (cb < bytes) {
      cb <<= 1;
      ++c;
    }
    if (c < kClasses) {
      std::lock_guard<std::mutex> l(Mu());
      if (Retained() + cb <= kMaxRetained) {
        Free()[c].push_back(p);
        Retained() += cb;
        return;
      }
    }
    free(p);
  }

 private:
  static constexpr int kClasses = 16;  // 64 KB .. 2 GB
  static std::mutex& Mu() {
    static std::mutex m;
    return m;
  }
  static std::vector<void*>* Free() {
    static std::vector<void*> v[kClasses];
    return v;
  }
  static size_t& Retained() {
    static size_t r = 0;
    return r;
  }
};

// Minimal vector of trivially copyable elements on BlockPool (std::vector with a custom allocator constructs element by
// element, which turns a 2 MB codestream copy into a byte loop).  Growth leaves new elements uninitialised unless a fill
// value is given.
template <class T>
class PodVec {
 public:
  PodVec() = default;
  PodVec(const PodVec& o) { assign(o.begin(), o.end()); }
  PodVec(PodVec&& o) noexcept : p_(o.p_), n_(o.n_), cap_(o.cap_) { o.p_ = nullptr; o.n_ = o.cap_ = 0; }
  PodVec& operator=(const PodVec& o) {
    if (this != &o) assign(o.begin(), o.end());
    return *this;
  }
  PodVec& operator=(PodVec&& o) noexcept {
    if (this != &o) {
      Release();
      p_ = o.p_; n_ = o.n_; cap_ = o.cap_;
      o.p_ = nullptr; o.n_ = o.cap_ = 0;
    }
    return *this;
  }
  ~PodVec() { Release(); }
  T* data() { return p_; }
  const T* data() const { return p_; }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  T* begin() { return p_; }
  T* end() { return p_ + n_; }
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }
  T& operator[](size_t i) { return p_[i]; }
  const T& operator[](size_t i) const { return p_[i]; }
  void clear() { n_ = 0; }
  void reserve(size_t n) {
    if (n <= cap_) return;
    size_t bytes = n * sizeof(T), cb = bytes;
    T* q;
    if (bytes >= BlockPool::kMinBytes) q = static_cast<T*>(BlockPool::Get(bytes, &cb));
    else if (!(q = static_cast<T*>(malloc(bytes)))) throw std::bad_alloc();
    if (n_) memcpy(q, p_, n_ * sizeof(T));
    const size_t keep = n_;
    Release();
    p_ = q;
    n_ = keep;
    cap_ = cb / sizeof(T);
  }
  template <class It>
  void assign(It first, It last) {
    n_ = 0;
    append(first, last);
  }
  template <class It>
  void append(It first, It last) {
    const size_t k = (size_t) (last - first);
    if (n_ + k > cap_) reserve(n_ + k > 2 * cap_ ? n_ + k : 2 * cap_);
    if (k) memcpy(p_ + n_, &*first, k * sizeof(T));
    n_ += k;
  }
  void resize(size_t n, T fill) {
    if (n > cap_) reserve(n);
    for (size_t i = n_; i < n; ++i) p_[i] = fill;
    n_ = n;
  }
  void resize(size_t n) { resize(n, T()); }

 private:
  void Release() {
    if (!p_) return;
    if (cap_ * sizeof(T) >= BlockPool::kMinBytes) BlockPool::Put(p_, cap_ * sizeof(T));
    else free(p_);
    p_ = nullptr;
    n_ = cap_ = 0;
  }
  T* p_ = nullptr;
  size_t n_ = 0, cap_ = 0;
};

using ByteVec = PodVec<uint8_t>;
using U16Vec = PodVec<uint16_t>;

}  // namespace jxlb
