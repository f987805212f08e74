/*
 * ref_graph_wrapper.cpp — C entry points around the REFERENCE's own GraphAdjacencySetUndirectedNoType (test infrastructure).
 * oracle/_ref/libref_graph.so = src/edgegraph3d/plgs/{graph_no_type,graph_adjacency_set_no_type,graph_adjacency_set_undirected_no_type}.cpp
 * compiled unmodified from /root/reference (see oracle/Makefile and oracle/ref_stubs/) + this file.  eg3d_ref_graph_link is the
 * statement convertEdgeImagePixelToGraph_NoCycles repeats for every neighbour pair
 * (src/edgegraph3d/io/input/convert_edge_images_pixel_to_segment.cpp:379-381): the bounded breadth-first loop check whose
 * `visited` flags survive between calls is the reference's compiled code here, not a restatement.
 */
#include "graph_adjacency_set_undirected_no_type.hpp"

extern "C" {
void* eg3d_ref_graph_new(unsigned long n) { return new GraphAdjacencySetUndirectedNoType(n); }
void eg3d_ref_graph_free(void* g) { delete (GraphAdjacencySetUndirectedNoType*)g; }
int eg3d_ref_graph_link(void* gp, unsigned long p, unsigned long c, unsigned long max_dist) {
  GraphAdjacencySetUndirectedNoType* g = (GraphAdjacencySetUndirectedNoType*)gp;
  if (p != c && !g->is_connected(p, c, max_dist)) { g->add_edge(p, c); return 1; }
  return 0;
}
void eg3d_ref_graph_add_edge(void* gp, unsigned long a, unsigned long b) { ((GraphAdjacencySetUndirectedNoType*)gp)->add_edge(a, b); }
unsigned long eg3d_ref_graph_degree(void* gp, unsigned long node) { return ((GraphAdjacencySetUndirectedNoType*)gp)->adjacency_lists[node].size(); }
void eg3d_ref_graph_neighbours(void* gp, unsigned long node, unsigned long* out) {
  for (unsigned long nb : ((GraphAdjacencySetUndirectedNoType*)gp)->adjacency_lists[node]) *out++ = nb;
}
/* get_components (graph_adjacency_set_undirected_no_type.cpp:44-69): component id per node, in the reference's discovery order */
void eg3d_ref_graph_components(void* gp, unsigned long* comp_of) {
  GraphAdjacencySetUndirectedNoType* g = (GraphAdjacencySetUndirectedNoType*)gp;
  vector<vector<ulong>> comps = g->get_components();
  for (unsigned long c = 0; c < comps.size(); c++) for (ulong n : comps[c]) comp_of[n] = c;
}
}
