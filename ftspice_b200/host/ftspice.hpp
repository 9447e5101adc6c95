// ftspice.hpp — C++ host side above the C-ABI (include/ftsb.h), mirroring the reference's front end and `Engine`.
//
// The reference is Rust and no Rust toolchain exists in the build image, so the host side is written in C++ with
// the reference's names, argument meaning and error behaviour (all citations relative to /root/reference/):
//   parse_spice_file / parse_value      src/parser.rs:17-62, 310-340   (grammar: src/spice.pest:1-59)
//   check_elems                         src/parser/check_elems.rs:6-22
//   NodeCollection                      src/node_collection.rs:13-94
//   Command                             src/command.rs:2-21
//   Engine::new / run_op / run_dc / run_tran   src/engine.rs:31-206
//   SimResult::push / get / print       src/engine/sim_result.rs:11-46
//   NotConvergedError                   src/engine/error.rs:3-10
// Nothing here computes circuit arithmetic: all of it happens in libftsb.so on the GPU.
#pragma once
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

struct ftsb_circuit;

namespace ftspice {

enum class Kind { R = 0, C = 1, L = 2, V = 3, I = 4, D = 5, Q = 6, M = 7 };
enum class FnKind { None = 0, Sin = 1, Pulse = 2, Exp = 3 };

// one element: the fields of the Rust device structs (src/device/*.rs); `nodes` AFTER the parser's swaps
struct Elem {
    Kind kind;
    std::string name;
    std::vector<std::string> nodes;
    double val = 0.0;
    FnKind fn_kind = FnKind::None;
    double fn[7] = {0, 0, 0, 0, 0, 0, 0};
};

struct Command {
    enum Type { Op, DC, Tran } type;
    std::string source; // DC
    double start = 0.0, stop = 0.0, step = 0.0;
};

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

double parse_value(const std::string &tok, const std::string &unit = "");
void parse_spice_text(const std::string &text, std::vector<Elem> &elems, std::vector<Command> &cmds);
void parse_spice_file(const std::string &path, std::vector<Elem> &elems, std::vector<Command> &cmds);
void check_elems(const std::vector<Elem> &elems); // throws "Duplicate elems found!" / "GND node not found!"

struct MNANode {
    bool is_current;
    int idx;
};

// name -> unknown index; iteration order = BTreeMap key order (bytewise)
class NodeCollection {
  public:
    static NodeCollection from_elems(const std::vector<Elem> &elems);
    static NodeCollection from_startup_elems(const std::vector<Elem> &elems);
    size_t len() const { return data_.size(); }
    std::optional<int> get_idx(const std::string &name) const;
    const std::map<std::string, MNANode> &items() const { return data_; }

  private:
    std::map<std::string, MNANode> data_;
};

struct NotConvergedError : std::runtime_error {
    NotConvergedError() : std::runtime_error("Couldn't converge simulation.") {}
};
// the reference process would have panicked (newtons_method.rs:53-55, nmos/model.rs:38-40)
struct EnginePanic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

std::string rust_f64_to_string(double x); // Rust's f64 Display: shortest round-trip, positional

class SimResult {
  public:
    explicit SimResult(const std::vector<std::string> &headers) : headers_(headers) {}
    void push(const std::map<std::string, double> &record);
    std::vector<double> get(const std::string &label) const;
    std::string to_csv() const;
    void print() const;
    const std::vector<std::string> &headers() const { return headers_; }
    const std::vector<std::vector<double>> &data() const { return data_; }

  private:
    std::vector<std::vector<double>> data_;
    std::vector<std::string> headers_;
};

class Engine {
  public:
    Engine(std::vector<Elem> elems, std::vector<Command> cmds, int device = 0);
    ~Engine();
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;

    std::optional<Command> op_cmd, dc_cmd, tran_cmd; // the first command of each kind (engine.rs:32-43)
    SimResult run_op();
    SimResult run_dc();
    SimResult run_tran();

    // index-lowered descriptor (what crosses the C-ABI), as JSON — used by the front-end tests
    std::string lowered_json() const;

  private:
    void ensure_handle();
    std::vector<Elem> elems_;
    NodeCollection run_nodes_, start_nodes_;
    int device_;
    ftsb_circuit *h_ = nullptr;
};

} // namespace ftspice
