#include "boost/numeric/ublas/matrix.hpp"
