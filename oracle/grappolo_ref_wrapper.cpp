/*
 * grappolo_ref_wrapper.cpp — C entry point for the REFERENCE's own community detection (test infrastructure).
 *
 * oracle/_ref/libgrappolo_ref.so = the vendored Grappolo sources compiled where they lie under
 * /root/reference/external/grappolo-05-2014 (see oracle/Makefile; no reference source is copied into this repository)
 * + this file, which fills clustering_parameters exactly as compute_communities does
 * (src/edgegraph3d/matching/polyline_matching/community_detection_interface.cpp:57-73) and calls
 * grappolo_community_detection (external/grappolo-05-2014/driverForGraphClustering_edited.cpp:48).
 * Only tests/ and profiles/ scripts load it, to compare the product's own Louvain with the reference's.
 */
#include <omp.h>
#include "defs.h"
#include "driverForGraphClustering_edited.hpp"

extern "C" int eg3d_ref_grappolo_communities(const char* graph_file, const char* communities_file, int n_threads) {
  if (n_threads > 0) omp_set_num_threads(n_threads);
  clustering_parameters opts;
  opts.inFile = graph_file;
  opts.ftype = 2;
  opts.coloring = true;
  opts.output = true;
  opts.minGraphSize = 10;
  return grappolo_community_detection(opts, communities_file);
}
