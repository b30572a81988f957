// Minimal stand-in for github.com/arloan/libimsux (absent here, unpinned in the reference) so that the
// reference's own headers compile UNMODIFIED for the oracle's reference-backed checker
// (oracle/_ref/libref_oip.so).  Only RAII / formatting / locking helpers -- no arithmetic lives here.
// TEST INFRASTRUCTURE ONLY.  Semantics inferred from the reference's call sites.
#pragma once
#include <cerrno>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>

#define IMSUX_USE_NS

template <typename T> struct imsux_default_dtor { void operator()(T *p) const { delete[] reinterpret_cast<char *>(p); } };
template <typename T> struct array_dtor { void operator()(T *p) const { delete[] p; } };
struct file_dtor { void operator()(FILE *f) const { if (f) fclose(f); } };

template <typename T, typename D = imsux_default_dtor<T>>
class scoped_ptr {
public:
    scoped_ptr() : p_(nullptr), d_() {}
    scoped_ptr(T *p) : p_(p), d_() {}
    scoped_ptr(std::nullptr_t) : p_(nullptr), d_() {}
    template <typename A> scoped_ptr(T *p, A a) : p_(p), d_(a) {}
    scoped_ptr(const scoped_ptr &) = delete;
    ~scoped_ptr() { reset(); }
    scoped_ptr &operator=(T *p) { attach(p); return *this; }
    scoped_ptr &operator=(const scoped_ptr &) = delete;
    operator T *() const { return p_; }
    T *operator->() const { return p_; }
    T *get() const { return p_; }
    bool is_null() const { return p_ == nullptr; }
    void attach(T *p) { reset(); p_ = p; }
    T *detach() { T *p = p_; p_ = nullptr; return p; }
private:
    void reset() { if (p_ && p_ != reinterpret_cast<T *>(-1)) d_(p_); p_ = nullptr; }
    T *p_;
    D d_;
};

template <typename T, typename D>
class scoped_ob {
public:
    scoped_ob(T v) : v_(v) {}
    ~scoped_ob() { D()(v_); }
    operator T() const { return v_; }
    T get() const { return v_; }
private:
    T v_;
};

struct errno_error : public std::runtime_error {
    explicit errno_error(const char *m) : std::runtime_error(std::string(m) + ": " + strerror(errno)) {}
    explicit errno_error(const std::string &m) : std::runtime_error(m + ": " + strerror(errno)) {}
};

struct xs {
    char s[2048];
    xs(const char *fmt, ...) {
        va_list ap; va_start(ap, fmt); vsnprintf(s, sizeof s, fmt, ap); va_end(ap);
    }
    operator const char *() const { return s; }
    operator std::string() const { return std::string(s); }
};

class comma_sep {
public:
    template <typename V> comma_sep(V v) { snprintf(b_, sizeof b_, "%.3f", (double)v); }
    const char *sep() const { return b_; }
private:
    char b_[64];
};

class stop_watch {
public:
    struct lap { double ellapsed; };
    stop_watch() : t0_(clock_::now()) {}
    lap tick() { auto t = clock_::now(); lap l{std::chrono::duration<double>(t - t0_).count()}; t0_ = t; return l; }
    static void rst() { g() = clock_::now(); }
    static lap tik() { return lap{std::chrono::duration<double>(clock_::now() - g()).count() + 1e-9}; }
private:
    typedef std::chrono::steady_clock clock_;
    static clock_::time_point &g() { static clock_::time_point t = clock_::now(); return t; }
    clock_::time_point t0_;
};

// critical section + scoped-lock block macro:  _ims_lock(CriticalSectionLocker, csl) { ... }
typedef std::recursive_mutex CRITICAL_SECTION;
inline void InitializeCriticalSection(CRITICAL_SECTION *) {}
class CriticalSectionLocker {
public:
    explicit CriticalSectionLocker(CRITICAL_SECTION &cs) : cs_(cs) {}
    void lock() { cs_.lock(); }
    void unlock() { cs_.unlock(); }
private:
    CRITICAL_SECTION &cs_;
};
template <typename L> struct imsux_scope_lock {
    L &l; bool once;
    explicit imsux_scope_lock(L &x) : l(x), once(true) { l.lock(); }
    ~imsux_scope_lock() { l.unlock(); }
};
#define _ims_lock(T, obj) for (imsux_scope_lock<T> _isl(obj); _isl.once; _isl.once = false)
