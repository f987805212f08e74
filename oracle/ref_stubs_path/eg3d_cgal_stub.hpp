/* Stand-in for the handful of CGAL types the reference's umbrella headers mention (TEST INFRASTRUCTURE).  The functions that use
 * them (3D line / plane compatibility of the legacy segment edge manager, geometric_utilities.cpp:1085-1368) are dead code on the
 * hot path; these definitions only let the translation unit compile and abort if anything geometric is ever asked of them. */
#pragma once
#include <cstdio>
#include <cstdlib>
#include "eg3d_boost_stub.hpp"
namespace CGAL {
[[noreturn]] inline void eg3d_cgal_unreachable() { fprintf(stderr, "eg3d CGAL stub called (off the hot path)\n"); abort(); }
template <typename FT> struct Cartesian;
template <typename K> struct Vector_3 {
  double x_, y_, z_;
  Vector_3(double a = 0, double b = 0, double c = 0) : x_(a), y_(b), z_(c) {}
  double squared_length() const { return x_ * x_ + y_ * y_ + z_ * z_; }
  double operator*(const Vector_3& o) const { return x_ * o.x_ + y_ * o.y_ + z_ * o.z_; }
  double x() const { return x_; } double y() const { return y_; } double z() const { return z_; }
};
template <typename K> struct Point_3 {
  double x_, y_, z_;
  Point_3(double a = 0, double b = 0, double c = 0) : x_(a), y_(b), z_(c) {}
  double x() const { return x_; } double y() const { return y_; } double z() const { return z_; }
  double operator[](int i) const { return i == 0 ? x_ : i == 1 ? y_ : z_; }
};
template <typename K> struct Line_3 {
  Point_3<K> a, b;
  Line_3() {}
  Line_3(const Point_3<K>& p, const Point_3<K>& q) : a(p), b(q) {}
  bool is_degenerate() const { return a.x_ == b.x_ && a.y_ == b.y_ && a.z_ == b.z_; }
  Vector_3<K> to_vector() const { return Vector_3<K>(b.x_ - a.x_, b.y_ - a.y_, b.z_ - a.z_); }
};
template <typename K> struct Plane_3 {
  Point_3<K> p, q, r;
  Plane_3() {}
  Plane_3(const Point_3<K>& a, const Point_3<K>& b, const Point_3<K>& c) : p(a), q(b), r(c) {}
  Vector_3<K> orthogonal_vector() const { eg3d_cgal_unreachable(); }
};
template <typename FT> struct Cartesian { typedef CGAL::Point_3<Cartesian> Point_3; typedef CGAL::Vector_3<Cartesian> Vector_3; typedef CGAL::Line_3<Cartesian> Line_3; typedef CGAL::Plane_3<Cartesian> Plane_3; };
template <typename FT> struct Simple_cartesian { typedef CGAL::Point_3<Simple_cartesian> Point_3; typedef CGAL::Vector_3<Simple_cartesian> Vector_3; };
template <typename K> boost::optional<boost::variant<Line_3<K>, Plane_3<K>>> intersection(const Plane_3<K>&, const Plane_3<K>&) { eg3d_cgal_unreachable(); }
}  // namespace CGAL
