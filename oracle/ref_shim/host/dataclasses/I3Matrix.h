// Stand-in for dataclasses/I3Matrix.h (a serializable boost::numeric::ublas::matrix<double>).
#ifndef CLSIM_REF_SHIM_I3MATRIX_H
#define CLSIM_REF_SHIM_I3MATRIX_H
#include "boost/numeric/ublas/matrix.hpp"
class I3Matrix : public boost::numeric::ublas::matrix<double> {
public:
    I3Matrix() {}
    I3Matrix(std::size_t rows, std::size_t cols) : boost::numeric::ublas::matrix<double>(rows, cols) {}
};
#endif
