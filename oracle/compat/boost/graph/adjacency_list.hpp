// compat shim (TEST INFRASTRUCTURE): the slice of Boost.Graph used by src/CostMatrix.cpp:98-109 —
// adjacency_list<vecS, vecS, undirectedS>, add_edge(u, v, g) growing the vertex set.
#ifndef RFS_COMPAT_BOOST_ADJ
#define RFS_COMPAT_BOOST_ADJ
#include <cstddef>
#include <vector>
namespace boost {
struct vecS {};
struct undirectedS {};
template <class OutEdgeS, class VertexS, class DirS>
class adjacency_list {
 public:
  std::vector<std::vector<std::size_t> > adj;
};
template <class A, class B, class C>
void add_edge(std::size_t u, std::size_t v, adjacency_list<A, B, C>& g) {
  const std::size_t n = (u > v ? u : v) + 1;
  if (g.adj.size() < n) g.adj.resize(n);
  g.adj[u].push_back(v);
  if (u != v) g.adj[v].push_back(u);
}
template <class A, class B, class C>
std::size_t num_vertices(const adjacency_list<A, B, C>& g) { return g.adj.size(); }
}
#endif
