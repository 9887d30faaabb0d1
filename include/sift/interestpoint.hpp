// sift::InterestPoint — the reference's result type (interestpoint.hpp:13-63), field for field.
#ifndef SIFT_INTERESTPOINT_HPP
#define SIFT_INTERESTPOINT_HPP

#include <vector>

#include "core.hpp"

namespace sift {

class InterestPoint {
   public:
    f32_t scale = 0;
    u16_t octave = 0;
    u16_t index = 0;
    bool filtered = false;
    Point<u16_t, u16_t> loc{0, 0};
    f32_t orientation = 0;
    std::vector<f32_t> descriptors;

    InterestPoint() = default;
    explicit InterestPoint(Point<u16_t, u16_t> at, f32_t s, u16_t oct, u16_t idx) : scale(s), octave(oct), index(idx), loc(at) {}

    // strict weak order "unfiltered before filtered" used by the reference's cleanup sort (sift.cpp:37,49)
    static bool cmpByFilter(const InterestPoint& a, const InterestPoint& b) { return !a.filtered && b.filtered; }
};

}  // namespace sift
#endif  // SIFT_INTERESTPOINT_HPP
