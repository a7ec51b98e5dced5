// TEST INFRASTRUCTURE — shadows the reference's controller/eigen.hpp (which needs <Eigen/Eigen>)
// with the typedefs the MPPI path, the forecasters and the Franka-Ridgeback objectives use.
#pragma once
#include <Eigen/Core>
using VectorXd = Eigen::VectorXd;
using MatrixXd = Eigen::MatrixXd;
using Vector2d = Eigen::Vector2d;
using Vector3d = Eigen::Vector3d;
using Vector4d = Eigen::Vector4d;
using Vector6d = Eigen::Vector6;
using Quaterniond = Eigen::Quaterniond;
using AngleAxisd = Eigen::AngleAxisd;
#ifndef M_PI
#define M_PI 3.141592653589793
#endif
