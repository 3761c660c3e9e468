// murty_compat.hpp — host side of the Murty compatibility switch (rfsb200_filter_cfg::murty_compat, quirk Q7).
//
// For a partition of the (eval point, measurement) graph with nR + nC > 8 the reference does not add up every
// assignment: it builds the (nR + nC) x (nR + nC) log-likelihood matrix [log L | diag log(1 - P_D); diag log(kappa) | 0],
// asks Murty's algorithm for the 200 best assignments that differ in their real part and adds up exp(score) of those
// (include/RBPHDFilter.hpp:904-959, src/MurtyAlgorithm.cpp:141-336, include/HungarianMethod.hpp).  The real part of an
// assignment is: every eval point is detected by one measurement of its own or missed, every measurement left over is
// clutter — so the sum asked for is the sum of the 200 LARGEST terms of the exact sum the device computes.
//
// k_best_sum() enumerates those terms in descending order with Murty's partitioning over a rectangular assignment
// problem (rows = eval points; columns = the measurements, then one "missed" column per row), each sub-problem solved by
// the shortest-augmenting-path form of the Hungarian method.  Written for this library; sizes are tiny (nR <= 32,
// nC <= 64, 200 x nR sub-problems).
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <queue>
#include <vector>

namespace rfsb200 {
namespace murty {

constexpr double BIG_NEG = -1000.0;   // the reference's log(0) (include/RBPHDFilter.hpp:869)
constexpr double FORBID = 1e200;      // cost of a forbidden cell

// min-cost assignment of n rows to m >= n columns (every row one column, no column twice); cost row-major n x m.
// Returns false if some row can only take forbidden cells.
inline bool assign_min(const std::vector<double>& cost, int n, int m, std::vector<int>& row_to_col, double& total) {
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double> u(n + 1, 0.0), v(m + 1, 0.0), minv(m + 1);
  std::vector<int> p(m + 1, 0), way(m + 1, 0);
  std::vector<char> used(m + 1);
  for (int i = 1; i <= n; i++) {
    p[0] = i;
    int j0 = 0;
    std::fill(minv.begin(), minv.end(), INF);
    std::fill(used.begin(), used.end(), 0);
    do {
      used[j0] = 1;
      const int i0 = p[j0];
      double delta = INF;
      int j1 = 0;
      for (int j = 1; j <= m; j++) {
        if (used[j]) continue;
        const double cur = cost[(size_t)(i0 - 1) * m + (j - 1)] - u[i0] - v[j];
        if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
        if (minv[j] < delta) { delta = minv[j]; j1 = j; }
      }
      if (!(delta < FORBID * 0.5)) return false;   // only forbidden cells left for this row
      for (int j = 0; j <= m; j++) {
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
        else minv[j] -= delta;
      }
      j0 = j1;
    } while (p[j0] != 0);
    do {
      const int j1 = way[j0];
      p[j0] = p[j1];
      j0 = j1;
    } while (j0);
  }
  row_to_col.assign(n, -1);
  total = 0;
  for (int j = 1; j <= m; j++)
    if (p[j]) {
      row_to_col[p[j] - 1] = j - 1;
      total += cost[(size_t)(p[j] - 1) * m + (j - 1)];
    }
  return total < FORBID * 0.5;
}

struct Node {
  double cost;                 // of the best assignment under the node's constraints (minimised)
  std::vector<int> assign;     // row -> column
  std::vector<double> c;       // the node's cost matrix (constraints written into it)
  int first_free;              // rows below are fixed
  bool operator<(const Node& o) const { return cost > o.cost; }   // priority_queue: smallest cost on top
};

// Sum of exp(score) over the (at most) k_max best real assignments of one partition, stopping like the reference when
// the assignments run out or a score drops below BIG_NEG.  L: nR x nC likelihoods (0 = no edge), pd: P_D per row,
// log_kappa: log clutter intensity (uniform).  n_terms (optional): how many terms were added.
inline double k_best_sum(const double* L, const double* pd, int nR, int nC, double log_kappa, int k_max, int* n_terms = nullptr) {
  const int m = nC + nR;
  // score of an assignment = sum_r log(1 - P_D_r) + nC log kappa + sum over the detections of (log L - log(1 - P_D_r) - log kappa)
  std::vector<double> miss(nR);
  double base = (double)nC * log_kappa;
  for (int r = 0; r < nR; r++) { miss[r] = std::log(1.0 - pd[r]); base += miss[r]; }
  std::vector<double> c0((size_t)nR * m, FORBID);
  for (int r = 0; r < nR; r++) {
    for (int q = 0; q < nC; q++) {
      double l = L[(size_t)r * nC + q];
      l = (l == 0.0) ? BIG_NEG : std::log(l);
      if (l < BIG_NEG) l = BIG_NEG;
      c0[(size_t)r * m + q] = -(l - miss[r] - log_kappa);
    }
    c0[(size_t)r * m + nC + r] = 0.0;   // missed
  }
  std::priority_queue<Node> pq;
  {
    Node root;
    root.c = c0;
    root.first_free = 0;
    if (!assign_min(root.c, nR, m, root.assign, root.cost)) { if (n_terms) *n_terms = 0; return 0.0; }
    pq.push(std::move(root));
  }
  double sum = 0;
  int terms = 0;
  while (terms < k_max && !pq.empty()) {
    Node nd = pq.top();
    pq.pop();
    const double score = base - nd.cost;
    if (score < BIG_NEG) break;
    sum += std::exp(score);
    terms++;
    // Murty's partition of the node's remaining solutions: child i keeps rows first_free .. i-1 as assigned and forbids
    // row i its column
    for (int i = nd.first_free; i < nR; i++) {
      Node ch;
      ch.c = nd.c;
      ch.first_free = i;
      for (int r = nd.first_free; r < i; r++) {   // fix row r to its column: forbid everything else in the row and the column
        const int col = nd.assign[r];
        for (int q = 0; q < m; q++) if (q != col) ch.c[(size_t)r * m + q] = FORBID;
        for (int rr = 0; rr < nR; rr++) if (rr != r) ch.c[(size_t)rr * m + col] = FORBID;
      }
      ch.c[(size_t)i * m + nd.assign[i]] = FORBID;
      ch.first_free = i;
      if (assign_min(ch.c, nR, m, ch.assign, ch.cost)) pq.push(std::move(ch));
    }
  }
  if (n_terms) *n_terms = terms;
  return sum;
}

}  // namespace murty
}  // namespace rfsb200
