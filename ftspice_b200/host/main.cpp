// main.cpp — command line with the reference's behaviour (src/main.rs:18-47): argv[1] = netlist; runs .op, .dc, .tran
// (the first command of each kind, in that order) and prints each result as CSV; a convergence failure ends the
// process like Rust's `main() -> Result<(), NotConvergedError>` (prints `Error: NotConvergedError`, exit code 1).
//   ftspice_b200 <netlist.sp>            run on the GPU through libftsb.so
//   ftspice_b200 --lower <netlist.sp>    parse + index lowering only (no GPU): prints the C-ABI descriptor as JSON
#include <cstdio>
#include <cstring>

#include "ftspice.hpp"

int main(int argc, char **argv)
{
    using namespace ftspice;
    bool lower_only = false;
    const char *file = nullptr;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--lower"))
            lower_only = true;
        else
            file = argv[i];
    }
    if (!file) {
        std::fprintf(stderr, "Insufficient arguments. Specify spice file to simulate.\n");
        return 101; // Rust panic exit code
    }
    try {
        std::vector<Elem> elems;
        std::vector<Command> cmds;
        parse_spice_file(file, elems, cmds);
        check_elems(elems);
        Engine engine(elems, cmds);
        if (lower_only) {
            std::puts(engine.lowered_json().c_str());
            return 0;
        }
        if (engine.op_cmd) engine.run_op().print();
        if (engine.dc_cmd) engine.run_dc().print();
        if (engine.tran_cmd) engine.run_tran().print();
    } catch (const NotConvergedError &) {
        std::fprintf(stderr, "Error: NotConvergedError\n");
        return 1;
    } catch (const EnginePanic &e) {
        std::fprintf(stderr, "thread 'main' panicked: %s\n", e.what());
        return 101;
    } catch (const ParseError &e) {
        std::fprintf(stderr, "thread 'main' panicked: %s\n", e.what());
        return 101;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    return 0;
}
