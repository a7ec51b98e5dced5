// The reference's CSV logging for the MPPI path with the same classes, file names, headers and number
// formatting (reference src/logging/file.hpp, logging/csv.hpp:64-170, logging/mppi.{hpp,cpp}), so runs of the
// device engine produce files the reference's analysis scripts read unchanged. With real Eigen on
// the include path the reference's own logging/*.hpp also compile against mppi_b200/trajectory.hpp (its getters
// have the reference's const signatures); this header is the self-contained equivalent.
#pragma once
#include <filesystem>
#include <fstream>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "mppi_b200/trajectory.hpp"

namespace logger {

class File {
public:
    static inline std::unique_ptr<File> create(std::filesystem::path path) {
        if (!path.parent_path().empty() && !std::filesystem::exists(path.parent_path())) {
            std::error_code code;
            if (!std::filesystem::create_directories(path.parent_path(), code)) {
                std::cerr << "failed to create log file " << path << ". " << code.message() << std::endl;
                return nullptr;
            }
        }
        std::fstream stream{path, std::ios::out};
        if (!stream.is_open()) { std::cerr << "failed to open log file " << path << std::endl; return nullptr; }
        return std::unique_ptr<File>(new File(std::move(stream)));
    }
    inline std::fstream &get_stream() { return m_stream; }
    template <typename T> inline void write(T &&value) { m_stream << value; }
    template <typename T> inline std::fstream &operator<<(T &&value) { m_stream << value; return m_stream; }
    inline ~File() { m_stream.flush(); m_stream.close(); }
private:
    explicit File(std::fstream &&out) : m_stream(std::move(out)) {}
    std::fstream m_stream;
};

class CSV {
public:
    using Header = std::vector<std::string>;
    struct Configuration {
        std::filesystem::path path;
        Header header;
    };
    template <typename... Args> static inline Header make_header(Args... args) {
        Header header;
        (push(header, args), ...);
        return header;
    }
    static inline std::unique_ptr<CSV> create(const Configuration &configuration) {
        auto file = File::create(configuration.path);
        if (!file) { std::cerr << "failed to create csv log file" << std::endl; return nullptr; }
        for (std::size_t i = 0; i < configuration.header.size(); i++) *file << (i ? ", " : "") << configuration.header[i];
        if (!configuration.header.empty()) *file << '\n';
        auto csv = std::unique_ptr<CSV>(new CSV());
        csv->m_file = std::move(file);
        return csv;
    }
    // one row; a std::vector<double> argument contributes every element (csv.hpp:150-163)
    template <typename Arg, typename... Args> void write(const Arg &arg, const Args &...args) {
        value(arg);
        ((*m_file << ", ", value(args)), ...);
        *m_file << '\n';
    }
    inline void flush() { m_file->get_stream().flush(); }
private:
    CSV() = default;
    static void push(Header &h, const std::string &s) { h.push_back(s); }
    static void push(Header &h, const char *s) { h.emplace_back(s); }
    static void push(Header &h, const std::vector<std::string> &v) { for (auto &s : v) h.push_back(s); }
    template <typename T> void value(const T &v) { *m_file << v; }
    void value(const std::vector<double> &v) { for (std::size_t i = 0; i < v.size(); i++) *m_file << (i ? ", " : "") << v[i]; }
    std::unique_ptr<File> m_file;
};

class MPPI {
public:
    struct Configuration {   // logging/mppi.hpp:17-54
        std::filesystem::path folder;
        unsigned int state_dof;
        unsigned int control_dof;
        std::size_t rollouts;
        bool log_costs = true, log_weights = true, log_gradient = true, log_optimal_rollout = true, log_optimal_cost = true, log_update = true;
    };

    static std::unique_ptr<MPPI> create(const Configuration &configuration) {   // logging/mppi.cpp:9-82
        std::vector<std::string> control, rollouts;
        for (unsigned int i = 1; i < configuration.control_dof + 1; i++) control.push_back("control" + std::to_string(i));
        for (std::size_t i = 1; i < configuration.rollouts + 1; i++) rollouts.push_back("rollout" + std::to_string(i));
        auto mppi = std::unique_ptr<MPPI>(new MPPI());
        auto open = [&](const char *name, CSV::Header header) { return CSV::create(CSV::Configuration{configuration.folder / name, std::move(header)}); };
        if (configuration.log_costs) mppi->m_costs = open("costs.csv", CSV::make_header("update", "time", rollouts));
        if (configuration.log_weights) mppi->m_weights = open("weights.csv", CSV::make_header("update", "time", rollouts));
        if (configuration.log_gradient) mppi->m_gradient = open("gradient.csv", CSV::make_header("update", "time", control));
        if (configuration.log_optimal_rollout) mppi->m_optimal_rollout = open("optimal_rollout.csv", CSV::make_header("update", "time", control));
        if (configuration.log_optimal_cost) mppi->m_optimal_cost = open("optimal_cost.csv", CSV::make_header("update", "time", "cost"));
        if (configuration.log_update) mppi->m_update = open("update.csv", CSV::make_header("update", "time", "update_duration"));
        const bool error = (configuration.log_costs && !mppi->m_costs) || (configuration.log_weights && !mppi->m_weights) ||
                           (configuration.log_gradient && !mppi->m_gradient) || (configuration.log_optimal_rollout && !mppi->m_optimal_rollout) ||
                           (configuration.log_optimal_cost && !mppi->m_optimal_cost) || (configuration.log_update && !mppi->m_update);
        if (error) { std::cerr << "failed to create csv logger" << std::endl; return nullptr; }
        mppi->m_last_update = std::numeric_limits<double>::min();
        return mppi;
    }

    void log(const mppi::Trajectory &trajectory) {   // logging/mppi.cpp:84-136
        const double time = trajectory.get_update_last();
        if (time == m_last_update) return;
        const double step = trajectory.get_time_step();
        const unsigned int steps = trajectory.get_step_count();
        const std::size_t iteration = trajectory.get_update_count();
        if (m_update) m_update->write(iteration, time, trajectory.get_update_duration());
        m_time.resize(steps);
        for (unsigned int i = 0; i < steps; ++i) m_time[i] = time + i * step;
        if (m_costs) {
            std::vector<double> costs;
            for (const auto &rollout : trajectory.get_rollouts()) costs.push_back(rollout.cost);
            m_costs->write(iteration, time, costs);
        }
        if (m_weights) {
            const auto &w = trajectory.get_weights();
            std::vector<double> weights((std::size_t)w.size());
            for (std::size_t k = 0; k < weights.size(); k++) weights[k] = w[(std::ptrdiff_t)k];
            m_weights->write(iteration, time, weights);
        }
        auto column = [](const MatrixXd &m, unsigned int i) { std::vector<double> c((std::size_t)m.rows()); for (std::size_t r = 0; r < c.size(); r++) c[r] = m((std::ptrdiff_t)r, i); return c; };
        if (m_gradient) for (unsigned int i = 0; i < steps; ++i) m_gradient->write(iteration, m_time[i], column(trajectory.get_gradient(), i));
        if (m_optimal_rollout) for (unsigned int i = 0; i < steps; ++i) m_optimal_rollout->write(iteration, m_time[i], column(trajectory.get_optimal_rollout(), i));
        if (m_optimal_cost) m_optimal_cost->write(iteration, time, trajectory.get_optimal_total_cost());
        m_last_update = time;
    }

private:
    MPPI() = default;
    double m_last_update;
    std::vector<double> m_time;
    std::unique_ptr<CSV> m_costs, m_weights, m_gradient, m_optimal_rollout, m_optimal_cost, m_update;
};

}  // namespace logger
