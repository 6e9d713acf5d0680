// TEST INFRASTRUCTURE. Force-included (-include) when compiling the reference sources with g++:
// utils.h:139-146 assigns std::chrono::high_resolution_clock::now() to a
// std::chrono::steady_clock::time_point, which only MSVC accepts (there the two clocks are one type).
#include <chrono>
#define high_resolution_clock steady_clock
