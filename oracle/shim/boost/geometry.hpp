// oracle/shim/boost/geometry.hpp -- TEST INFRASTRUCTURE, not product code.
//
// Stand-in for the piece of Boost.Geometry that NuWisdom.h / NuWisdom.cpp use: a 2-D cartesian point, the euclidean distance
// and an "r-tree" with insert / size / clear / iteration / k-nearest query -- kept as a plain vector with a linear-scan KNN
// (same answers as an r-tree up to the order of equidistant points).  Boost is not in this image.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <iterator>
#include <utility>
#include <vector>

namespace boost { namespace geometry {

namespace cs { struct cartesian {}; }
namespace model {
template <class T, std::size_t N, class CS> class point {
public:
    point() { for (std::size_t i = 0; i < N; ++i) v[i] = T(); }
    point(T a, T b) { v[0] = a; v[1] = b; }
    template <std::size_t K> T get() const { return v[K]; }
    template <std::size_t K> void set(T x) { v[K] = x; }
private:
    T v[N];
};
}  // namespace model

template <class T, std::size_t N, class CS>
inline double distance(const model::point<T, N, CS> &a, const model::point<T, N, CS> &b) {
    const double dx = a.template get<0>() - b.template get<0>(), dy = a.template get<1>() - b.template get<1>();
    return std::sqrt(dx * dx + dy * dy);
}

namespace index {
template <std::size_t M> struct quadratic {};
template <class P> struct ax_nearest { P target; unsigned k; };
template <class P> inline ax_nearest<P> nearest(const P &p, unsigned k) { return ax_nearest<P>{p, k}; }

template <class Value, class Params> class rtree {
public:
    typedef typename std::vector<Value>::const_iterator const_iterator;
    void insert(const Value &v) { mValues.push_back(v); }
    std::size_t size() const { return mValues.size(); }
    void clear() { mValues.clear(); }
    const_iterator begin() const { return mValues.begin(); }
    const_iterator end() const { return mValues.end(); }
    template <class P, class Out> std::size_t query(const ax_nearest<P> &q, Out out) const {
        std::vector<std::pair<double, std::size_t>> d(mValues.size());
        for (std::size_t i = 0; i < mValues.size(); ++i) d[i] = std::make_pair(distance(q.target, mValues[i].first), i);
        const std::size_t k = std::min<std::size_t>(q.k, d.size());
        std::partial_sort(d.begin(), d.begin() + k, d.end());
        for (std::size_t i = 0; i < k; ++i) *out++ = mValues[d[i].second];
        return k;
    }
private:
    std::vector<Value> mValues;
};
}  // namespace index

}}  // namespace boost::geometry
