// config.hpp -- reader for the reference's controller configuration files.
//
// The reference parses config/*.yaml with yaml-cpp (src/main.cu:455-628).  yaml-cpp is not a
// dependency here: the schema is flat (scalars, lists of scalars, one nested map `cost`), so a
// small indentation-aware reader covers it, in both block ("- 1") and flow ("[1, 2]") list
// syntax.  Keys and required-key behaviour follow parse_config: a missing key prints the
// reference's message and exits with status 1.  verify() restates verify_parse
// (src/main.cu:686-725), the reference's only config test.
#ifndef MPPI_CPP_CONFIG_HPP_
#define MPPI_CPP_CONFIG_HPP_

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace mppi_cfg {

struct Config {
    std::string env;            // `env`      (model file of the plant)
    int samples = 0;            // `samples`
    int state_dim = 0;          // `state-dim`
    int act_dim = 0;            // `action-dim`
    int horizon = 0;            // `horizon`
    float dt = 0;               // `dt`
    float lambda = 0;           // `lambda`   (parsed and dropped by the reference, :311)
    std::vector<float> noise;   // `noise`    (idem)
    std::vector<float> init_act;  // `init-act` (idem)
    std::vector<float> max_a;   // `max-a`    (idem)
    std::vector<float> goal;    // `goal`
    std::string cost_type;      // `cost.type`
    std::vector<float> cost_w;  // `cost.w`
};

namespace detail {

inline std::string trim(const std::string &s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

inline std::string strip_comment(const std::string &s)
{
    size_t p = s.find('#');
    return p == std::string::npos ? s : s.substr(0, p);
}

inline std::vector<std::string> split_flow(const std::string &v)
{
    std::vector<std::string> out;
    std::string body = v.substr(1, v.size() - 2), item;
    std::stringstream ss(body);
    while (std::getline(ss, item, ',')) {
        item = trim(item);
        if (!item.empty()) out.push_back(item);
    }
    return out;
}

// flat store: "key" -> scalar, "key" -> list, nested as "cost.w"
struct Doc {
    std::map<std::string, std::string> scalars;
    std::map<std::string, std::vector<std::string>> lists;
    bool has(const std::string &k) const { return scalars.count(k) || lists.count(k); }
};

inline Doc load(const std::string &path)
{
    std::ifstream in(path);
    if (!in) {
        std::printf("Cannot open config file %s\n", path.c_str());
        std::exit(1);
    }
    Doc d;
    std::string line, parent, list_key;
    int parent_indent = -1, base_indent = -1;
    while (std::getline(in, line)) {
        line = strip_comment(line);
        if (trim(line).empty() || trim(line) == "---" || trim(line) == "...") continue;
        int indent = (int)line.find_first_not_of(' ');
        std::string t = trim(line);
        if (base_indent < 0) base_indent = indent;
        if (t[0] == '-') {                       // block list item
            if (!list_key.empty()) d.lists[list_key].push_back(trim(t.substr(1)));
            continue;
        }
        size_t colon = t.find(':');
        if (colon == std::string::npos) continue;
        std::string key = trim(t.substr(0, colon)), val = trim(t.substr(colon + 1));
        if (indent <= parent_indent || indent == base_indent) { parent.clear(); parent_indent = -1; }
        std::string full = parent.empty() ? key : parent + "." + key;
        if (val.empty()) {                       // map or block list follows
            list_key = full;
            d.lists[full];                       // tentatively a list; a child key makes it a map
            if (parent.empty()) { parent = key; parent_indent = indent; }
        } else if (val.front() == '[' && val.back() == ']') {
            d.lists[full] = split_flow(val);
            list_key.clear();
        } else {
            if (val.size() >= 2 && (val.front() == '"' || val.front() == '\''))
                val = val.substr(1, val.size() - 2);
            d.scalars[full] = val;
            list_key.clear();
        }
    }
    return d;
}

inline void require(const Doc &d, const std::string &key, const char *msg)
{
    if (!d.has(key) || (d.lists.count(key) && d.lists.at(key).empty() && !d.scalars.count(key))) {
        // a map parent (e.g. `cost`) is stored as an empty list: look for children
        bool child = false;
        for (auto &kv : d.scalars) child |= kv.first.rfind(key + ".", 0) == 0;
        for (auto &kv : d.lists) child |= kv.first.rfind(key + ".", 0) == 0;
        if (!child) {
            std::printf("%s\n", msg);
            std::exit(1);
        }
    }
}

inline std::vector<float> floats(const Doc &d, const std::string &key)
{
    std::vector<float> v;
    auto it = d.lists.find(key);
    if (it != d.lists.end())
        for (auto &s : it->second) v.push_back(std::strtof(s.c_str(), nullptr));
    return v;
}

}  // namespace detail

// parse_config, src/main.cu:455-628 (same required keys, same messages)
inline Config parse(const std::string &path)
{
    using namespace detail;
    Doc d = load(path);
    Config c;
    require(d, "env", "Please provide a env file in the config file");
    c.env = d.scalars["env"];
    require(d, "samples", "Please provide the number of samples in the config file");
    c.samples = std::atoi(d.scalars["samples"].c_str());
    require(d, "state-dim", "Please provide the state dimension in the config file");
    c.state_dim = std::atoi(d.scalars["state-dim"].c_str());
    require(d, "action-dim", "Please provide the action dimension in the config file");
    c.act_dim = std::atoi(d.scalars["action-dim"].c_str());
    require(d, "horizon", "Please provide the prediction horizon in the config file");
    c.horizon = std::atoi(d.scalars["horizon"].c_str());
    require(d, "dt", "Please provide the time step in the config file");
    c.dt = std::strtof(d.scalars["dt"].c_str(), nullptr);
    require(d, "lambda", "Please provide a env file in the config file");   // sic, :520
    c.lambda = std::strtof(d.scalars["lambda"].c_str(), nullptr);
    require(d, "noise", "Please provide a noise vector in the config file, should be a array of size action-dim");
    c.noise = floats(d, "noise");
    require(d, "init-act", "Please provide a init vector in the config file, should be a array of size action-dim");
    c.init_act = floats(d, "init-act");
    require(d, "max-a", "Please provide a max input vector in the config file, should be a array of size action-dim");
    c.max_a = floats(d, "max-a");
    if ((int)c.max_a.size() != c.act_dim)
        std::printf("Warning: the input limit is different than the action dimension \n");
    require(d, "goal", "Please provide a goal vector in the config file, should be a array of size action-dim");
    c.goal = floats(d, "goal");
    if ((int)c.goal.size() != c.state_dim)
        std::printf("Warning: the goal size is different than the state dimension \n");
    require(d, "cost", "Please provide cost function in the config file.");
    require(d, "cost.type", "Please provide cost function type in the config file. Currently supported: quadratic ");
    c.cost_type = d.scalars["cost.type"];
    require(d, "cost.w", "Please provide cost function type in the config file. Currently supported: quadratic ");
    c.cost_w = floats(d, "cost.w");
    if ((int)c.cost_w.size() != c.state_dim)
        std::printf("Warning: the cost function weights matrix is different than the state dimension \n");
    return c;
}

// verify_parse, src/main.cu:686-725: the constants of config/mppi-config-test.yaml, TOL 1e-6
inline bool verify_test_config(const Config &c)
{
    const double tol = 1e-6;
    auto eq = [&](double a, double b) { return std::fabs(a - b) < tol; };
    bool ok = c.samples == 3 && c.state_dim == 4 && c.act_dim == 2 && c.horizon == 12;
    ok = ok && eq(c.dt, 0.1) && eq(c.lambda, 1.5);
    ok = ok && c.max_a.size() == 2 && eq(c.max_a[0], 1.2) && eq(c.max_a[1], 1.3);
    ok = ok && c.noise.size() == 2 && eq(c.noise[0], 0.24) && eq(c.noise[1], 0.26);
    ok = ok && c.init_act.size() == 2 && eq(c.init_act[0], 0.1) && eq(c.init_act[1], 0.2);
    ok = ok && c.cost_w.size() == 4 && eq(c.cost_w[0], 1) && eq(c.cost_w[1], 2) &&
         eq(c.cost_w[2], 0.5) && eq(c.cost_w[3], 0.75);
    ok = ok && c.goal.size() == 4 && eq(c.goal[0], 1) && eq(c.goal[1], 2) && eq(c.goal[2], 3) &&
         eq(c.goal[3], 4);
    return ok;
}

}  // namespace mppi_cfg
#endif
