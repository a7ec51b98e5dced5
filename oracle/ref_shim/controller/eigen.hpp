// TEST INFRASTRUCTURE — shadows the reference's controller/eigen.hpp (which needs <Eigen/Eigen>)
// with the two typedefs the MPPI path uses.
#pragma once
#include <Eigen/Core>
using VectorXd = Eigen::VectorXd;
using MatrixXd = Eigen::MatrixXd;
using Vector6d = Eigen::Vector6;
