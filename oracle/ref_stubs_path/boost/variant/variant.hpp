/* stand-in */
#include "../../eg3d_boost_stub.hpp"
