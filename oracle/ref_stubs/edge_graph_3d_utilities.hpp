/* Stand-in for the reference's umbrella header include/edgegraph3d/utils/edge_graph_3d_utilities.hpp (test infrastructure).
 * That header drags in OpenCV and CGAL, which this image does not have; the reference's graph classes
 * (src/edgegraph3d/plgs/graph_no_type.cpp, graph_adjacency_set_no_type.cpp, graph_adjacency_set_undirected_no_type.cpp) only
 * need the standard containers and `ulong` from it.  With this directory first on the include path those three source files
 * compile UNMODIFIED, where they lie (oracle/Makefile, target _ref/libref_graph.so). */
#pragma once
#include <sys/types.h>
#include <iostream>
#include <map>
#include <set>
#include <stack>
#include <vector>
using namespace std;
