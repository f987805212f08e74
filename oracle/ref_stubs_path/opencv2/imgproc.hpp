/* stand-in: see ../eg3d_cv_stub.hpp */
#include "../eg3d_cv_stub.hpp"
