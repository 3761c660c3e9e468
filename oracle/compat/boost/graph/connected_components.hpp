// compat shim (TEST INFRASTRUCTURE): boost::connected_components.  Boost runs a depth-first search
// from every still-unvisited vertex in ascending vertex order and numbers the components in the
// order they are discovered, i.e. by their lowest vertex — the property the reference's
// partition labelling (and quirk Q6) depends on.
#ifndef RFS_COMPAT_BOOST_CC
#define RFS_COMPAT_BOOST_CC
#include <vector>
#include "adjacency_list.hpp"
namespace boost {
template <class A, class B, class C, class ComponentMap>
int connected_components(const adjacency_list<A, B, C>& g, ComponentMap comp) {
  const std::size_t n = g.adj.size();
  std::vector<char> seen(n, 0);
  int ncc = 0;
  std::vector<std::size_t> stack;
  for (std::size_t s = 0; s < n; s++) {
    if (seen[s]) continue;
    stack.assign(1, s);
    seen[s] = 1;
    while (!stack.empty()) {
      std::size_t u = stack.back();
      stack.pop_back();
      comp[u] = ncc;
      for (std::size_t k = 0; k < g.adj[u].size(); k++) {
        std::size_t v = g.adj[u][k];
        if (!seen[v]) { seen[v] = 1; stack.push_back(v); }
      }
    }
    ncc++;
  }
  return ncc;
}
}
#endif
