/*
 * dsref_gsl.h -- TEST INFRASTRUCTURE (oracle/_ref build only).  The slice of Microsoft GSL (ms-gsl, Dependencies.md:9, not
 * vendored) that the reference's sources use: narrow_cast, narrow, span, make_span.
 */
#ifndef DSREF_GSL_H
#define DSREF_GSL_H
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace gsl {
template <class T, class U> constexpr T narrow_cast(U&& u) noexcept { return static_cast<T>(std::forward<U>(u)); }

struct narrowing_error : public std::exception {
};
template <class T, class U> T narrow(U u)
{
    T t = narrow_cast<T>(u);
    if (static_cast<U>(t) != u) throw narrowing_error();
    if ((t < T{}) != (u < U{})) throw narrowing_error();
    return t;
}

template <class T> class span {
public:
    span() : p(nullptr), n(0) {}
    span(T* ptr, std::ptrdiff_t count) : p(ptr), n(count) {}
    template <class A> span(std::vector<A>& v) : p(v.data()), n((std::ptrdiff_t)v.size()) {}
    T& operator[](std::ptrdiff_t i) const { return p[i]; }
    std::ptrdiff_t size() const { return n; }
    T* data() const { return p; }
    T* begin() const { return p; }
    T* end() const { return p + n; }

private:
    T* p;
    std::ptrdiff_t n;
};
template <class T> span<T> make_span(T* ptr, std::ptrdiff_t count) { return span<T>(ptr, count); }
template <class T> span<T> make_span(std::vector<T>& v) { return span<T>(v.data(), (std::ptrdiff_t)v.size()); }
} // namespace gsl
#endif
