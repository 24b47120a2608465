// Stress test of the host thread pool of svs_ba_set_problem (no GPU): sums, nested begin/end, idle gaps.
#include <cstdio>
#include <chrono>
#include <numeric>
#include <thread>
#include <vector>
#include "../../scavislam_b200/csrc/host_pool.hpp"
int main() {
  svs::SpinPool pool(4);
  long long bad = 0;
  for (int call = 0; call < 200; ++call) {
    pool.begin();
    for (int rep = 0; rep < 50; ++rep) {
      const int n = 1 + (call * 7 + rep * 13) % 97;
      std::vector<long long> out(n, 0);
      std::vector<int> hits(n, 0);
      pool.parallel_for(n, [&](int i) {
        long long s = 0;
        for (int k = 0; k <= i * 100; ++k) s += k;
        out[i] = s;
        hits[i]++;
      });
      for (int i = 0; i < n; ++i) {
        const long long m = (long long)i * 100;
        if (out[i] != m * (m + 1) / 2 || hits[i] != 1) ++bad;
      }
    }
    pool.end();
    if (call % 50 == 0) std::this_thread::sleep_for(std::chrono::milliseconds(5));
    // a loop outside begin()/end() runs inline
    int cnt = 0;
    pool.parallel_for(5, [&](int) { ++cnt; });
    if (cnt != 5) ++bad;
  }
  printf("%s\n", bad ? "FAIL" : "OK");
  return bad != 0;
}
