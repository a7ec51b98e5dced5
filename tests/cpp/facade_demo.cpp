// Drives the C++ facade (mppi::Trajectory over the C ABI) the way the reference's Actor does
// (src/simulation/frankaridgeback/actor.cpp:96-101,187-190,201) and prints results for the test to
// compare with the oracle. Usage: facade_demo <toy|track|assisted> <K> <horison> <updates> <noise.bin|-> <out.bin>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

#include "mppi_b200/logging.hpp"
#include "mppi_b200/systems.hpp"

static const double kPi = 3.14159265358979323846;

int main(int argc, char **argv) {
    if (argc < 7) return 2;
    const std::string which = argv[1];
    const long K = std::atol(argv[2]);
    const double horison = std::atof(argv[3]);
    const int updates = std::atoi(argv[4]);
    mppi::Configuration c;
    // optional 7th.. arguments: keep_best, smoothing (0/1), cadence, x0 values for the toy, "refrng"
    const long keep_best = argc > 7 ? std::atol(argv[7]) : 20;
    const bool smoothing = argc > 8 ? std::atoi(argv[8]) != 0 : true;
    const double cadence = argc > 9 ? std::atof(argv[9]) : 0.05;
    const bool refrng = argc > 10 && std::strcmp(argv[10], "refrng") == 0;
    c.rollouts = K; c.keep_best_rollouts = keep_best; c.time_step = 0.01; c.horison = horison; c.gradient_step = 2.0; c.cost_scale = 10.0; c.cost_discount_factor = 1.0;
    c.control_bound = true; c.threads = 1;
    if (smoothing) c.smoothing = mppi::Configuration::Smoothing{(unsigned)(std::getenv("MPPI_B200_DEMO_SG_WINDOW") ? std::atoi(std::getenv("MPPI_B200_DEMO_SG_WINDOW")) : 10), 1};
    std::unique_ptr<mppi::Dynamics> dyn; std::unique_ptr<mppi::Cost> cost;
    VectorXd x0;
    if (which == "toy") {
        c.covariance = MatrixXd(2, 2); c.covariance(0, 0) = 1; c.covariance(1, 1) = 1;
        c.control_min = VectorXd(2); c.control_max = VectorXd(2);
        for (int i = 0; i < 2; i++) { c.control_min[i] = -5; c.control_max[i] = 5; }
        c.initial_state = VectorXd(4);
        dyn = mppi_b200::DoubleIntegrator::create();
        cost = mppi_b200::PointCost::create(mppi_b200::PointCost::default_configuration());
        x0 = VectorXd(4);
        for (int i = 0; i < 4 && 11 + i < argc; i++) x0[i] = std::atof(argv[11 + i]);
    } else {
        const double var[12] = {0.1, 0.1, 0.2, 7.5, 7.5, 7.5, 7.5, 7.5, 7.5, 7.5, 0.0, 0.0};           // base.hpp:79-83
        const double lim[12] = {0.5, 0.5, 1.0, 100, 100, 100, 100, 100, 100, 100, 0.05, 0.05};          // base.hpp:85-94
        c.covariance = MatrixXd(12, 12); c.control_min = VectorXd(12); c.control_max = VectorXd(12);
        for (int i = 0; i < 12; i++) { c.covariance(i, i) = var[i]; c.control_min[i] = -lim[i]; c.control_max[i] = lim[i]; }
        c.control_default = VectorXd(12);
        c.initial_state = VectorXd(31);
        x0 = VectorXd(31);
        const double q[12] = {0.2, 0.2, kPi / 4, 0.0, kPi / 5, 0.0, -kPi / 2, 0.0, 2, kPi / 4, 0.025, 0.025};  // state.cpp:15-19
        for (int i = 0; i < 12; i++) x0[i] = q[i];
        x0[30] = argc > 11 ? std::atof(argv[11]) : 10.0;
        if (which == "track") {
            dyn = FrankaRidgeback::PinocchioDynamics::create();
            cost = FrankaRidgeback::TrackPoint::create(FrankaRidgeback::TrackPoint::default_configuration());
        } else if (which == "assisted_locf") {
            // the wrench forecast as a device producer: last observation carried forward (forecast.hpp:62-140)
            LOCFForecast::Configuration lc; lc.observation = VectorXd(6); lc.horison = 1e9;
            std::shared_ptr<Forecast> forecast = LOCFForecast::create(lc);
            if (!forecast) { std::fprintf(stderr, "forecast create failed\n"); return 3; }
            VectorXd w(6); w[0] = 10.0;
            forecast->update(w, 0.0);
            dyn = FrankaRidgeback::PinocchioDynamics::create(FrankaRidgeback::PinocchioDynamics::default_configuration(), forecast);
            auto p = FrankaRidgeback::AssistedManipulation::default_configuration();
            p.enable_energy_limit = 1; p.link_position_mode = MPPI_B200_LINKS_BODY_COM;
            cost = FrankaRidgeback::AssistedManipulation::create(p);
        } else {
            dyn = FrankaRidgeback::PinocchioDynamics::create(FrankaRidgeback::PinocchioDynamics::default_configuration(), [](double) { return std::array<double, 6>{10.0, 0, 0, 0, 0, 0}; });
            auto p = FrankaRidgeback::AssistedManipulation::default_configuration();
            p.enable_energy_limit = 1; p.link_position_mode = MPPI_B200_LINKS_BODY_COM;
            cost = FrankaRidgeback::AssistedManipulation::create(p);
        }
    }
    auto trajectory = mppi::Trajectory::create(c, std::move(dyn), std::move(cost));
    if (!trajectory) return 3;
    trajectory->use_reference_rng(refrng);
    const std::size_t nu = trajectory->get_control_dof(), T = trajectory->get_step_count(), R = trajectory->get_rollout_count();
    std::vector<double> noise;
    if (std::strcmp(argv[5], "-") != 0) {
        std::ifstream f(argv[5], std::ios::binary);
        noise.resize((std::size_t)updates * R * T * nu);
        f.read(reinterpret_cast<char *>(noise.data()), (std::streamsize)(noise.size() * 8));
        if (!f) return 4;
    }
    std::ofstream out(argv[6], std::ios::binary);
    // the reference's CSV logger (logging/mppi.cpp) over the device trajectory, as BaseTest does (test/case/base.cpp:52-63)
    std::unique_ptr<logger::MPPI> mppi_logger;
    if (const char *folder = std::getenv("MPPI_B200_DEMO_LOG")) {
        logger::MPPI::Configuration lc;
        lc.folder = folder; lc.state_dof = trajectory->get_state_dof(); lc.control_dof = trajectory->get_control_dof(); lc.rollouts = trajectory->get_rollout_count();
        mppi_logger = logger::MPPI::create(lc);
        if (!mppi_logger) return 5;
    }
    for (int u = 0; u < updates; u++) {
        if (!noise.empty()) trajectory->set_injected_noise(noise.data() + (std::size_t)u * R * T * nu);
        trajectory->update(x0, cadence * u);
        const MatrixXd &U = trajectory->trajectory();
        out.write(reinterpret_cast<const char *>(U.data()), (std::streamsize)(nu * T * 8));
        VectorXd ctl = (*trajectory)(cadence * u + 0.013);
        out.write(reinterpret_cast<const char *>(ctl.data()), (std::streamsize)(nu * 8));
        const double oc = trajectory->get_optimal_total_cost();
        out.write(reinterpret_cast<const char *>(&oc), 8);
        const auto &w = trajectory->get_weights();
        out.write(reinterpret_cast<const char *>(w.data()), (std::streamsize)(R * 8));
        if (mppi_logger) { mppi_logger->log(*trajectory); mppi_logger->log(*trajectory); }   // the second call writes nothing
    }
    mppi_logger.reset();
    if (which == "assisted" || which == "assisted_locf") {
        const auto &am = dynamic_cast<const FrankaRidgeback::AssistedManipulation &>(trajectory->get_optimal_cost());
        std::printf("breakdown %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", am.get_joint_limit_cost(), am.get_self_collision_cost(), am.get_workspace_cost(),
                    am.get_energy_tank_cost(), am.get_joint_velocity_cost(), am.get_trajectory_cost(), am.get_manipulability_cost());
    }
    // error conventions: an unknown pair is refused at create (no CPU fallback)
    struct Foreign : mppi::Cost {
        std::unique_ptr<mppi::Cost> copy() override { return nullptr; }
        void reset(double) override {}
        double get_cost(const VectorXd &, const VectorXd &, mppi::Dynamics *, double) override { return 0; }
        int get_control_dof() override { return 2; }
        int get_state_dof() override { return 4; }
    };
    auto refused = mppi::Trajectory::create(c, mppi_b200::DoubleIntegrator::create(), std::make_unique<Foreign>());
    std::printf("foreign_cost_refused %d\n", refused == nullptr);
    std::printf("updates %zu rollouts %zu steps %zu duration %.3e\n", trajectory->get_update_count(), R, T, trajectory->get_update_duration());
    return 0;
}
