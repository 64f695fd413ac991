// Stand-in for the sliver of boost::numeric::ublas the reference's I3CLSimVectorTransformMatrix uses: a dense row-major
// matrix, a vector, prod(matrix, vector) and noalias().
#ifndef CLSIM_REF_SHIM_UBLAS_MATRIX_HPP
#define CLSIM_REF_SHIM_UBLAS_MATRIX_HPP
#include <cstddef>
#include <vector>
namespace boost { namespace numeric { namespace ublas {
template <class T> class vector {
public:
    vector() {}
    explicit vector(std::size_t n) : v_(n, T()) {}
    std::size_t size() const { return v_.size(); }
    T &operator()(std::size_t i) { return v_[i]; }
    const T &operator()(std::size_t i) const { return v_[i]; }
    T &operator[](std::size_t i) { return v_[i]; }
    const T &operator[](std::size_t i) const { return v_[i]; }
    typename std::vector<T>::iterator begin() { return v_.begin(); }
    typename std::vector<T>::iterator end() { return v_.end(); }
    typename std::vector<T>::const_iterator begin() const { return v_.begin(); }
    typename std::vector<T>::const_iterator end() const { return v_.end(); }
private:
    std::vector<T> v_;
};
template <class T> class matrix {
public:
    matrix() : r_(0), c_(0) {}
    matrix(std::size_t r, std::size_t c) : r_(r), c_(c), v_(r * c, T()) {}
    std::size_t size1() const { return r_; }
    std::size_t size2() const { return c_; }
    T &operator()(std::size_t i, std::size_t j) { return v_[i * c_ + j]; }
    const T &operator()(std::size_t i, std::size_t j) const { return v_[i * c_ + j]; }
private:
    std::size_t r_, c_;
    std::vector<T> v_;
};
// row by row, left to right: the order ublas' dense row-major prod() accumulates in
template <class T> inline vector<T> prod(const matrix<T> &m, const vector<T> &x)
{
    vector<T> out(m.size1());
    for (std::size_t i = 0; i < m.size1(); ++i) {
        T sum = T();
        for (std::size_t j = 0; j < m.size2(); ++j) sum += m(i, j) * x(j);
        out(i) = sum;
    }
    return out;
}
template <class V> inline V &noalias(V &v) { return v; }
}}} // namespace boost::numeric::ublas
#endif
