// Minimal column-major VectorXd / MatrixXd so the facade compiles where Eigen is not installed
// (the build container). With Eigen on the include path the real types are used and the facade's
// signatures are the reference's verbatim (src/controller/mppi.hpp:321-470).
#pragma once
#if defined(MPPI_B200_USE_EIGEN) || (__has_include(<Eigen/Core>) && !defined(MPPI_B200_NO_EIGEN))
#include <Eigen/Core>
using VectorXd = Eigen::VectorXd;
using MatrixXd = Eigen::MatrixXd;
namespace mppi_b200 { template <class T> using Ref = Eigen::Ref<T>; }
#else
#include <cstddef>
#include <vector>
namespace mppi_b200 {
class MatrixXd {
public:
    MatrixXd() = default;
    MatrixXd(std::ptrdiff_t rows, std::ptrdiff_t cols) : m_rows(rows), m_cols(cols), m_data((std::size_t)rows * cols, 0.0) {}
    static MatrixXd Zero(std::ptrdiff_t r, std::ptrdiff_t c) { return MatrixXd(r, c); }
    std::ptrdiff_t rows() const { return m_rows; }
    std::ptrdiff_t cols() const { return m_cols; }
    std::ptrdiff_t size() const { return m_rows * m_cols; }
    double *data() { return m_data.data(); }
    const double *data() const { return m_data.data(); }
    double &operator()(std::ptrdiff_t r, std::ptrdiff_t c) { return m_data[(std::size_t)(r + m_rows * c)]; }
    double operator()(std::ptrdiff_t r, std::ptrdiff_t c) const { return m_data[(std::size_t)(r + m_rows * c)]; }
    void setZero() { for (auto &x : m_data) x = 0.0; }
    void resize(std::ptrdiff_t r, std::ptrdiff_t c) { m_rows = r; m_cols = c; m_data.assign((std::size_t)r * c, 0.0); }
protected:
    std::ptrdiff_t m_rows = 0, m_cols = 0;
    std::vector<double> m_data;
};
class VectorXd : public MatrixXd {
public:
    VectorXd() = default;
    explicit VectorXd(std::ptrdiff_t n) : MatrixXd(n, 1) {}
    static VectorXd Zero(std::ptrdiff_t n) { return VectorXd(n); }
    double &operator[](std::ptrdiff_t i) { return m_data[(std::size_t)i]; }
    double operator[](std::ptrdiff_t i) const { return m_data[(std::size_t)i]; }
    double &operator()(std::ptrdiff_t i) { return m_data[(std::size_t)i]; }
    double operator()(std::ptrdiff_t i) const { return m_data[(std::size_t)i]; }
    void resize(std::ptrdiff_t n) { MatrixXd::resize(n, 1); }
};
template <class T> using Ref = T &;
}  // namespace mppi_b200
using VectorXd = mppi_b200::VectorXd;
using MatrixXd = mppi_b200::MatrixXd;
#endif
