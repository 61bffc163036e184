// plan.h - host-side planner: racu_desc -> canonical kernel parameters (see plan.cc).
#pragma once
#include <string>
#include "kcore.h"
#include "static_progs.h"

namespace racu {

constexpr int RACU_ERR_PEEL = 100; // internal: peel the outermost expression axis on the host and retry
constexpr int RACU_ERR_SEQ = 101;  // internal: a fold that must narrow to the destination at every step (integer destination, floating
                                   // point right hand side): run the folded axes one index at a time, in row-major order


struct Plan
{
    bool empty = false;          // nothing to traverse
    bool fill_identity = false;  // empty reduction in overwrite mode: store the identity
    KParams kp;
    unsigned grid_x = 0, grid_y = 1;
    size_t smem = 0;
    bool need_finish = false;
    FinishParams fp;
    size_t partial_bytes = 0;
    const char * kernel = "";
    bool is_static = false;      // a compile-time specialised kernel matched (static_kernels.cu)
    SParams sp;
    SProg jit_prog {};           // sp.id == SP_JIT: the program the kernel is specialised for at run time (jit.cc)
    void * jit_fn = nullptr;     // its loaded kernel (CUfunction)
};

// Set by jit.cc when the library carries the run-time specialiser (never in the CPU plan simulator).
extern bool (*g_jit_prepare)(Plan &, std::string &);
// Set by static_kernels.cu: CTAs of the plan's specialised kernel that fit on one SM (0 when there is no device to ask).
extern int (*g_static_occupancy)(Plan const &);

// What the expression computes, before any layout decision.
struct Shape { int rank; int64_t len[KMAXR]; };

int check_program(racu_desc const * d, std::string & err);                 // result dtype or -1
int agree(racu_desc const * d, Shape & s, std::string & err);             // racu_status
// reduce_out != nullptr: whole-expression reduction into *reduce_out (device, dtype = program result type).
int make_plan(racu_desc const * d, int redop, void * reduce_out, Plan & plan, std::string & err);

void peel_axis0(racu_desc const & d, int64_t i, racu_desc & out);
void peel_axis(racu_desc const & d, int k, int64_t i, racu_desc & out);  // same for any expression axis k
int first_folded_axis(racu_desc const * d, Shape const & sh);           // first axis the destination does not move along, or -1
// Structural checks that must hold before anything dereferences the descriptor by index: operand / instruction counts,
// destination index, operand kinds / dtypes / ranks, and the program (check_program). racu_status.
int validate_desc(racu_desc const * d, bool reducing, std::string & err);
// True when the traversal may be split along expression axis 0 into independent launches run in sequence.
bool can_peel(racu_desc const * d, bool reducing);
bool peel_last_only(racu_desc const * d);
bool peel_reduction(racu_desc const & d, int redop, void * reduce_out, racu_desc & out, int & acc_type, uint64_t & identity);

size_t dtype_size(int t);

} // namespace racu
