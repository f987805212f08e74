/* Stand-in for the Boost pieces the reference's headers name (TEST INFRASTRUCTURE): serialization hooks (member templates that
 * are never instantiated here), optional / variant (dead CGAL code) and filesystem (file output off the hot path). */
#pragma once
#include <string>
#include <cstdio>
#include <cstdlib>
namespace boost {
namespace serialization {
class access;
template <typename B, typename D> B& base_object(D& d) { return static_cast<B&>(d); }
}  // namespace serialization
namespace archive {
struct text_oarchive { template <typename S> explicit text_oarchive(S&) {} template <typename T> text_oarchive& operator<<(const T&) { fprintf(stderr, "eg3d boost stub: serialization is not modelled\n"); abort(); } template <typename T> text_oarchive& operator&(const T& t) { return *this << t; } };
struct text_iarchive { template <typename S> explicit text_iarchive(S&) {} template <typename T> text_iarchive& operator>>(T&) { fprintf(stderr, "eg3d boost stub: serialization is not modelled\n"); abort(); } template <typename T> text_iarchive& operator&(T& t) { return *this >> t; } };
}  // namespace archive
template <typename T> struct optional {
  bool has; T v;
  optional() : has(false), v() {}
  optional(const T& t) : has(true), v(t) {}
  explicit operator bool() const { return has; }
  T& operator*() { return v; }
  const T& operator*() const { return v; }
};
template <typename A, typename B> struct variant { int which_; A a; B b; variant() : which_(0) {} variant(const A& x) : which_(0), a(x) {} variant(const B& x) : which_(1), b(x) {} };
template <typename T, typename A, typename B> const T* get(const variant<A, B>*) { return nullptr; }
template <typename T, typename A, typename B> T* get(variant<A, B>*) { return nullptr; }
namespace filesystem {
struct path { std::string s; path() {} path(const std::string& x) : s(x) {} path(const char* x) : s(x) {} std::string string() const { return s; } const char* c_str() const { return s.c_str(); } path parent_path() const { return path(); } path filename() const { return *this; } path stem() const { return *this; } };
inline bool exists(const path&) { fprintf(stderr, "eg3d boost stub: filesystem is not modelled\n"); abort(); }
inline bool is_directory(const path&) { abort(); }
inline bool is_regular_file(const path&) { abort(); }
inline bool create_directory(const path&) { abort(); }
inline bool create_directories(const path&) { abort(); }
inline std::string basename(const path& p) { return p.s; }
inline std::string extension(const path& p) { return p.s; }
struct directory_iterator { directory_iterator() {} explicit directory_iterator(const path&) { abort(); } bool operator!=(const directory_iterator&) const { return false; } directory_iterator& operator++() { return *this; } path operator*() const { return path(); } const directory_iterator* operator->() const { return this; } path path_() const { return path(); } };
}  // namespace filesystem
}  // namespace boost
