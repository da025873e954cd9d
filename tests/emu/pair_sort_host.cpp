// TEST INFRASTRUCTURE: spgemm_b200/csrc/pair_sort.h (the device code k_s1_heavy calls) compiled as plain C++.
#include "../../spgemm_b200/csrc/pair_sort.h"
extern "C" void host_pair_heap_sort(int *ka, int *kb, int n) { tsg::pair_heap_sort(ka, kb, n); }
