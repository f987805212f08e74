/* stand-in */
#include "../eg3d_cgal_stub.hpp"
