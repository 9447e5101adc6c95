// ftspice.cpp — see ftspice.hpp.  Citations relative to /root/reference/.
#include "ftspice.hpp"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>

#include "../../include/ftsb.h"

namespace ftspice {

// ---------------------------------------------------------------------------------------------------------
// values: number = "-"? ("0" | [1-9][0-9]*) ("." [0-9]*)? ; prefix in grammar order (src/spice.pest:44-46);
// value *= mult in f64 (src/parser.rs:321-336)
// ---------------------------------------------------------------------------------------------------------
static bool eq_nocase(const std::string &a, const std::string &b)
{
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i)
        if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
    return true;
}

double parse_value(const std::string &tok, const std::string &unit)
{
    size_t i = 0;
    if (i < tok.size() && tok[i] == '-') ++i;
    if (i >= tok.size() || !std::isdigit((unsigned char)tok[i])) throw ParseError("bad value " + tok);
    if (tok[i] == '0') {
        ++i;
    } else {
        while (i < tok.size() && std::isdigit((unsigned char)tok[i])) ++i;
    }
    if (i < tok.size() && tok[i] == '.') {
        ++i;
        while (i < tok.size() && std::isdigit((unsigned char)tok[i])) ++i;
    }
    std::string num = tok.substr(0, i);
    if (!num.empty() && num.back() == '.') num += "0";
    double val = std::strtod(num.c_str(), nullptr);
    static const std::pair<const char *, double> prefixes[] = {
        {"G", 1e9}, {"M", 1e6}, {"k", 1e3}, {"h", 1e2}, {"da", 1e1}, {"d", 1e-1},
        {"c", 1e-2}, {"m", 1e-3}, {"u", 1e-6}, {"n", 1e-9}, {"p", 1e-12}, {"f", 1e-15}};
    for (const auto &p : prefixes) {
        size_t l = std::strlen(p.first);
        if (tok.compare(i, l, p.first) == 0) {
            val *= p.second;
            i += l;
            break;
        }
    }
    const std::string rest = tok.substr(i);
    if (unit.empty() ? !rest.empty() : !eq_nocase(rest, unit)) throw ParseError("bad value " + tok);
    return val;
}

static std::vector<std::string> split_ws(const std::string &s)
{
    std::vector<std::string> out;
    std::string cur;
    for (char ch : s) {
        if (ch == ' ') {
            if (!cur.empty()) out.push_back(cur), cur.clear();
        } else {
            cur.push_back(ch);
        }
    }
    if (!cur.empty()) out.push_back(cur);
    return out;
}

// SpiceFn::eval at t = 0 (the parser stores it as the source value, src/parser.rs:95-99); the general evaluation
// is device code (src/spice_fn.rs:8-124)
static double spice_fn_t0(FnKind k, const double *p)
{
    const double t = 0.0;
    switch (k) {
    case FnKind::Sin:
        return p[0] + p[1] * std::sin(2.0 * 3.14159265358979323846 * p[2] * t);
    case FnKind::Pulse: {
        const double tn = std::fabs(std::fmod(t, p[6]));
        if (tn < p[2]) return p[0];
        if (tn < p[2] + p[3]) return p[0] + (tn - p[2]) / p[3] * (p[1] - p[0]);
        if (tn < p[2] + p[5] - p[4]) return p[1];
        if (tn < p[2] + p[5]) return p[1] + (tn - (p[2] + p[5] - p[4])) / p[4] * (p[0] - p[1]);
        return p[0];
    }
    case FnKind::Exp:
        if (t < p[2]) return p[0];
        if (t < p[4]) return p[0] + (p[0] - p[1]) * std::expm1(-(t - p[2]) / p[3]);
        return p[0] + (p[0] - p[1]) * std::expm1(-(t - p[2]) / p[3]) + (p[1] - p[0]) * std::expm1(-(t - p[4]) / p[5]);
    default:
        return 0.0;
    }
}

static void parse_fn(const std::string &text, Elem &e)
{
    const size_t lp = text.find('('), rp = text.rfind(')');
    if (lp == std::string::npos || rp == std::string::npos || rp < lp) throw ParseError("bad function " + text);
    const std::string name = text.substr(0, lp);
    size_t need = 0;
    if (eq_nocase(name, "SIN"))
        e.fn_kind = FnKind::Sin, need = 3;
    else if (eq_nocase(name, "PULSE"))
        e.fn_kind = FnKind::Pulse, need = 7;
    else if (eq_nocase(name, "EXP"))
        e.fn_kind = FnKind::Exp, need = 6;
    else
        throw ParseError("bad function " + text);
    const auto vals = split_ws(text.substr(lp + 1, rp - lp - 1));
    if (vals.size() != need) throw ParseError("wrong number of function values: " + text);
    for (size_t i = 0; i < need; ++i) e.fn[i] = parse_value(vals[i]);
    e.val = spice_fn_t0(e.fn_kind, e.fn);
}

void parse_spice_text(const std::string &text, std::vector<Elem> &elems, std::vector<Command> &cmds)
{
    std::istringstream in(text);
    std::string raw;
    bool ended = false;
    while (std::getline(in, raw)) {
        size_t a = raw.find_first_not_of(' '), b = raw.find_last_not_of(" \r");
        if (a == std::string::npos) continue;
        const std::string line = raw.substr(a, b - a + 1);
        if (line[0] == '*' || line[0] == '$') continue; // src/spice.pest:54-59
        std::string low = line;
        for (char &ch : low) ch = (char)std::tolower((unsigned char)ch);
        if (low.rfind(".end", 0) == 0) {
            ended = true;
            break;
        }
        const auto tk = split_ws(line);
        if (low == ".op") {
            cmds.push_back({Command::Op, "", 0, 0, 0});
            continue;
        }
        if (low.rfind(".dc", 0) == 0) { // src/parser.rs:225-239
            if (tk.size() != 5) throw ParseError("bad .dc: " + line);
            cmds.push_back({Command::DC, tk[1], parse_value(tk[2]), parse_value(tk[3]), parse_value(tk[4])});
            continue;
        }
        if (low.rfind(".tran", 0) == 0) { // src/parser.rs:241-252: start is always 0
            if (tk.size() != 3) throw ParseError("bad .tran: " + line);
            cmds.push_back({Command::Tran, "", 0.0, parse_value(tk[1]), parse_value(tk[2])});
            continue;
        }
        Elem e;
        e.name = tk[0];
        switch (std::toupper((unsigned char)line[0])) {
        case 'R':
        case 'C':
        case 'L': {
            if (tk.size() != 4) throw ParseError("bad element: " + line);
            const char letter = (char)std::toupper((unsigned char)line[0]);
            const size_t eq = tk[3].find('=');
            if (eq != 1 || std::toupper((unsigned char)tk[3][0]) != letter) throw ParseError("bad element value: " + line);
            e.val = parse_value(tk[3].substr(2));
            if (letter == 'R') {
                e.kind = Kind::R;
                e.nodes = {tk[1], tk[2]}; // parser.rs:67-69
            } else {
                e.kind = letter == 'C' ? Kind::C : Kind::L;
                e.nodes = {tk[2], tk[1]}; // parser.rs:145-146, 161-162: [second, first]
            }
            break;
        }
        case 'V':
        case 'I': {
            if (tk.size() < 4) throw ParseError("bad source: " + line);
            e.kind = std::toupper((unsigned char)line[0]) == 'V' ? Kind::V : Kind::I;
            e.nodes = {tk[2], tk[1]}; // parser.rs:83-84, 114-115
            std::string rest;
            for (size_t i = 3; i < tk.size(); ++i) rest += (i > 3 ? " " : "") + tk[i];
            if (rest.find('(') != std::string::npos)
                parse_fn(rest, e);
            else
                e.val = parse_value(rest, e.kind == Kind::V ? "V" : "A");
            break;
        }
        case 'D':
            if (tk.size() != 4 || tk[3] != "d_model") throw ParseError("bad diode: " + line);
            e.kind = Kind::D;
            e.nodes = {tk[2], tk[1]}; // parser.rs:177-178
            break;
        case 'Q':
        case 'M': {
            const bool q = std::toupper((unsigned char)line[0]) == 'Q';
            if (tk.size() != 6 || tk[5] != (q ? "q_model" : "t_model")) throw ParseError("bad transistor: " + line);
            e.kind = q ? Kind::Q : Kind::M;
            e.nodes = {tk[1], tk[2], tk[3]}; // the 4th netlist node is ignored (parser.rs:186-219)
            break;
        }
        default:
            throw ParseError("Unsuccessful parse: " + line);
        }
        elems.push_back(e);
    }
    if (!ended) throw ParseError("Unsuccessful parse: missing .end");
}

void parse_spice_file(const std::string &path, std::vector<Elem> &elems, std::vector<Command> &cmds)
{
    std::ifstream f(path);
    if (!f) throw ParseError("Cannot read file.");
    std::stringstream ss;
    ss << f.rdbuf();
    parse_spice_text(ss.str(), elems, cmds);
}

void check_elems(const std::vector<Elem> &elems)
{
    std::set<std::string> names;
    for (const auto &e : elems) names.insert(e.name);
    if (names.size() != elems.size()) throw ParseError("Duplicate elems found!");
    bool gnd = false;
    for (const auto &e : elems)
        for (const auto &n : e.nodes) gnd |= (n == "0");
    if (!gnd) throw ParseError("GND node not found!");
}

// ---------------------------------------------------------------------------------------------------------
// NodeCollection (src/node_collection.rs:13-74): std::set / std::map order strings bytewise like BTreeSet/BTreeMap
// ---------------------------------------------------------------------------------------------------------
NodeCollection NodeCollection::from_elems(const std::vector<Elem> &elems)
{
    NodeCollection nc;
    std::set<std::string> v_names, i_names;
    for (const auto &e : elems)
        for (const auto &n : e.nodes)
            if (n != "0") v_names.insert(n);
    int i = 0;
    for (const auto &n : v_names) nc.data_[n] = MNANode{false, i++};
    for (const auto &e : elems)
        if (e.kind == Kind::V) i_names.insert(e.name); // GType::G2 (vdd.rs:24-26)
    int j = 0;
    for (const auto &n : i_names) nc.data_[n] = MNANode{true, (int)v_names.size() + j++};
    return nc;
}

NodeCollection NodeCollection::from_startup_elems(const std::vector<Elem> &elems)
{
    NodeCollection nc = from_elems(elems);
    const int base = (int)nc.data_.size();
    std::set<std::string> i_names;
    for (const auto &e : elems)
        if (e.kind == Kind::L) i_names.insert(e.name); // (G1, G2) elements: inductors (ind.rs:26-32)
    int j = 0;
    for (const auto &n : i_names) nc.data_[n] = MNANode{true, base + j++};
    return nc;
}

std::optional<int> NodeCollection::get_idx(const std::string &name) const
{
    auto it = data_.find(name);
    if (it == data_.end()) return std::nullopt;
    return it->second.idx;
}

// ---------------------------------------------------------------------------------------------------------
// SimResult (src/engine/sim_result.rs:11-46)
// ---------------------------------------------------------------------------------------------------------
std::string rust_f64_to_string(double x)
{
    if (std::isnan(x)) return "NaN";
    if (std::isinf(x)) return x > 0 ? "inf" : "-inf";
    char buf[512];
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::fixed); // shortest round-trip, positional
    return std::string(buf, r.ptr);
}

void SimResult::push(const std::map<std::string, double> &record)
{
    std::vector<double> row;
    for (const auto &h : headers_) row.push_back(record.at(h));
    data_.push_back(row);
}

std::vector<double> SimResult::get(const std::string &label) const
{
    size_t idx = 0;
    for (; idx < headers_.size(); ++idx)
        if (headers_[idx] == label) break;
    if (idx == headers_.size()) throw std::runtime_error("Label not found");
    std::vector<double> out;
    for (const auto &row : data_) out.push_back(row[idx]);
    return out;
}

std::string SimResult::to_csv() const
{
    std::string out;
    for (size_t i = 0; i < headers_.size(); ++i) out += (i ? "," : "") + headers_[i];
    out += "\n";
    for (const auto &row : data_) {
        for (size_t i = 0; i < row.size(); ++i) out += (i ? "," : "") + rust_f64_to_string(row[i]);
        out += "\n";
    }
    return out;
}

void SimResult::print() const { std::fputs(to_csv().c_str(), stdout); }

// ---------------------------------------------------------------------------------------------------------
// Engine (src/engine.rs:22-206)
// ---------------------------------------------------------------------------------------------------------
static void raise_for(int status)
{
    if (status == FTSB_OK) return;
    if (status == FTSB_NOT_CONVERGED || status == FTSB_TRAN_H_UNDERFLOW) throw NotConvergedError();
    if (status == FTSB_DIVERGED) throw EnginePanic("v_err or i_err diverged");
    throw EnginePanic("internal error: entered unreachable code");
}

Engine::Engine(std::vector<Elem> elems, std::vector<Command> cmds, int device) : elems_(std::move(elems)), device_(device)
{
    for (const auto &c : cmds) { // first command of each kind (engine.rs:32-43)
        if (c.type == Command::Op && !op_cmd) op_cmd = c;
        if (c.type == Command::DC && !dc_cmd) dc_cmd = c;
        if (c.type == Command::Tran && !tran_cmd) tran_cmd = c;
    }
    run_nodes_ = NodeCollection::from_elems(elems_);
    start_nodes_ = NodeCollection::from_startup_elems(elems_);
}

Engine::~Engine()
{
    if (h_) ftsb_destroy(h_);
}

static int idx_or_gnd(const NodeCollection &nc, const std::string &name)
{
    if (name == "0") return -1;
    return nc.get_idx(name).value();
}

static void lower(const std::vector<Elem> &elems, const NodeCollection &start, std::vector<ftsb_elem> &out,
                  std::vector<int32_t> &cls, std::vector<double> &vals, std::vector<double> &fns)
{
    cls.assign(start.len(), 0);
    for (const auto &kv : start.items()) cls[kv.second.idx] = kv.second.is_current ? FTSB_ROW_CURRENT : FTSB_ROW_VOLTAGE;
    for (const auto &e : elems) {
        ftsb_elem fe;
        fe.kind = (int)e.kind;
        fe.node[0] = fe.node[1] = fe.node[2] = -1;
        for (size_t j = 0; j < e.nodes.size() && j < 3; ++j) fe.node[j] = idx_or_gnd(start, e.nodes[j]);
        fe.branch = (e.kind == Kind::V || e.kind == Kind::L) ? start.get_idx(e.name).value() : -1;
        fe.fn_kind = (int)e.fn_kind;
        out.push_back(fe);
        vals.push_back(e.val);
        if (e.fn_kind != FnKind::None) fns.insert(fns.end(), e.fn, e.fn + 7);
    }
}

void Engine::ensure_handle()
{
    if (h_) return;
    std::vector<ftsb_elem> fe;
    std::vector<int32_t> cls;
    std::vector<double> vals, fns;
    lower(elems_, start_nodes_, fe, cls, vals, fns);
    int rc = ftsb_compile(fe.data(), (int32_t)fe.size(), (int32_t)run_nodes_.len(), (int32_t)start_nodes_.len(), cls.data(),
                          device_, &h_);
    if (rc) throw std::runtime_error(std::string("ftsb_compile: ") + ftsb_last_error(nullptr));
    rc = ftsb_set_params(h_, 1, vals.data(), fns.empty() ? nullptr : fns.data(), FTSB_LAYOUT_INSTANCE_MAJOR, FTSB_MEM_HOST);
    if (rc) throw std::runtime_error(std::string("ftsb_set_params: ") + ftsb_last_error(h_));
}

std::string Engine::lowered_json() const
{
    std::vector<ftsb_elem> fe;
    std::vector<int32_t> cls;
    std::vector<double> vals, fns;
    lower(elems_, start_nodes_, fe, cls, vals, fns);
    std::ostringstream o;
    o << "{\"n_run\":" << run_nodes_.len() << ",\"n_start\":" << start_nodes_.len() << ",\"names\":[";
    std::vector<std::string> names(start_nodes_.len());
    for (const auto &kv : start_nodes_.items()) names[kv.second.idx] = kv.first;
    for (size_t i = 0; i < names.size(); ++i) o << (i ? "," : "") << "\"" << names[i] << "\"";
    o << "],\"cls\":[";
    for (size_t i = 0; i < cls.size(); ++i) o << (i ? "," : "") << cls[i];
    o << "],\"headers_start\":[";
    bool first = true;
    for (const auto &kv : start_nodes_.items()) o << (first ? "" : ",") << "\"" << kv.first << "\"", first = false;
    o << "],\"headers_run\":[";
    first = true;
    for (const auto &kv : run_nodes_.items()) o << (first ? "" : ",") << "\"" << kv.first << "\"", first = false;
    o << "],\"elems\":[";
    for (size_t i = 0; i < fe.size(); ++i) {
        o << (i ? "," : "") << "{\"name\":\"" << elems_[i].name << "\",\"kind\":" << fe[i].kind << ",\"node\":[" << fe[i].node[0]
          << "," << fe[i].node[1] << "," << fe[i].node[2] << "],\"branch\":" << fe[i].branch << ",\"fn_kind\":" << fe[i].fn_kind
          << ",\"val\":\"" << rust_f64_to_string(vals[i]) << "\",\"fn\":[";
        for (int j = 0; j < 7; ++j) o << (j ? "," : "") << "\"" << rust_f64_to_string(elems_[i].fn[j]) << "\"";
        o << "]}";
    }
    o << "],\"cmds\":[";
    first = true;
    auto put = [&](const std::optional<Command> &c, const char *k) {
        if (!c) return;
        o << (first ? "" : ",") << "{\"kind\":\"" << k << "\",\"source\":\"" << c->source << "\",\"start\":\""
          << rust_f64_to_string(c->start) << "\",\"stop\":\"" << rust_f64_to_string(c->stop) << "\",\"step\":\""
          << rust_f64_to_string(c->step) << "\"}";
        first = false;
    };
    put(op_cmd, "op");
    put(dc_cmd, "dc");
    put(tran_cmd, "tran");
    o << "]}";
    return o.str();
}

SimResult Engine::run_op()
{
    ensure_handle();
    const int n = (int)start_nodes_.len();
    std::vector<double> x(std::max(n, 1));
    int32_t it = 0, st = 0;
    if (ftsb_run_op(h_, x.data(), &it, &st, FTSB_MEM_HOST)) throw std::runtime_error(ftsb_last_error(h_));
    raise_for(st);
    std::vector<std::string> headers{"n_iters"};
    for (const auto &kv : start_nodes_.items()) headers.push_back(kv.first);
    SimResult res(headers);
    std::map<std::string, double> rec{{"n_iters", (double)it}};
    for (const auto &kv : start_nodes_.items()) rec[kv.first] = x[kv.second.idx];
    res.push(rec);
    return res;
}

SimResult Engine::run_dc()
{
    if (!dc_cmd) throw std::runtime_error("DC simulation wrongly configured.");
    int sweep = -1;
    for (size_t i = 0; i < elems_.size(); ++i)
        if (elems_[i].name == dc_cmd->source) {
            sweep = (int)i;
            break;
        }
    if (sweep < 0) throw std::runtime_error("Sweep source not found");
    ensure_handle();
    const int n = (int)run_nodes_.len();
    const double cnt = std::ceil((dc_cmd->stop - dc_cmd->start) / dc_cmd->step); // ndarray Array::range (engine.rs:120)
    const int64_t P = cnt >= 1.0 ? (int64_t)cnt : 1;
    std::vector<double> x((size_t)P * std::max(n, 1));
    std::vector<int32_t> it(P);
    int64_t done = 0;
    int32_t st = 0;
    if (ftsb_run_dc(h_, sweep, &dc_cmd->start, &dc_cmd->stop, &dc_cmd->step, P, x.data(), it.data(), &done, &st, FTSB_MEM_HOST))
        throw std::runtime_error(ftsb_last_error(h_));
    raise_for(st); // `?` on the first failing point: the reference prints nothing (engine.rs:127)
    std::vector<std::string> headers{"n_iters"};
    for (const auto &kv : run_nodes_.items()) headers.push_back(kv.first);
    SimResult res(headers);
    for (int64_t k = 0; k < done; ++k) {
        std::map<std::string, double> rec{{"n_iters", (double)it[k]}};
        for (const auto &kv : run_nodes_.items()) rec[kv.first] = x[(size_t)k * n + kv.second.idx];
        res.push(rec);
    }
    return res;
}

SimResult Engine::run_tran()
{
    if (!tran_cmd) throw std::runtime_error("DC simulation wrongly configured."); // sic (engine.rs:145)
    ensure_handle();
    const int n = (int)run_nodes_.len();
    int64_t cap = 4096;
    for (;;) {
        std::vector<double> t(cap), x((size_t)cap * std::max(n, 1));
        std::vector<int32_t> it(cap);
        int64_t nrec = 0;
        int32_t st = 0;
        ftsb_tran_opts opts{cap, 1, 0};
        if (ftsb_run_tran(h_, tran_cmd->stop, tran_cmd->step, &opts, t.data(), x.data(), it.data(), &nrec, nullptr, &st,
                          FTSB_MEM_HOST))
            throw std::runtime_error(ftsb_last_error(h_));
        raise_for(st);
        if (nrec > cap) { // the reference keeps every accepted step: retry with room for all of them
            cap = nrec + 16;
            continue;
        }
        std::vector<std::string> headers{"n_iters", "t"};
        for (const auto &kv : run_nodes_.items()) headers.push_back(kv.first);
        SimResult res(headers);
        for (int64_t k = 0; k < nrec; ++k) {
            std::map<std::string, double> rec{{"n_iters", (double)it[k]}, {"t", t[k]}};
            for (const auto &kv : run_nodes_.items()) rec[kv.first] = x[(size_t)k * n + kv.second.idx];
            res.push(rec);
        }
        return res;
    }
}

} // namespace ftspice
