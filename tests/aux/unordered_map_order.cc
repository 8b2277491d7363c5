#include <unordered_map>
#include <cstdio>
#include <vector>
#include <random>
#include <algorithm>
int main() {
  std::mt19937 g(1);
  int bad = 0;
  for (int trial = 0; trial < 2000; ++trial) {
    std::vector<unsigned> keys = {0,1,2,3,4,5,6};
    std::shuffle(keys.begin(), keys.end(), g);
    int k = 1 + (int)(g() % 7);
    std::unordered_map<unsigned, double> m;
    std::vector<unsigned> ins;
    for (int i = 0; i < k; ++i) { m[keys[i]] = i; ins.push_back(keys[i]); }
    std::vector<unsigned> it;
    for (auto& kv : m) it.push_back(kv.first);
    std::reverse(ins.begin(), ins.end());
    if (it != ins) { ++bad; if (bad < 4) { printf("insertion(rev):"); for (auto x : ins) printf(" %u", x); printf("  iteration:"); for (auto x : it) printf(" %u", x); printf("  buckets %zu\n", m.bucket_count()); } }
  }
  std::unordered_map<unsigned, double> m; m[3] = 1;
  printf("mismatches %d of 2000; bucket_count after 1 insert: %zu\n", bad, m.bucket_count());
}
