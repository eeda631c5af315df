// TEST INFRASTRUCTURE -- not part of the product path.
//
// C shim over the REFERENCE's own qpOASES 3.2 (sources stay under
// /root/reference/src/qpOASES; nothing is copied into this repo).  Compiled by
// oracle/Makefile together with those sources into oracle/_ref/libqpoases_ref.so.
//
// It issues exactly the call sequence of the reference hot path,
//   /root/reference/src/MPC_Ctrl/SolverMPC.cpp:529-539
//     QProblem problem_red(new_vars, new_cons); Options op; op.setToMPC();
//     op.printLevel = PL_NONE; problem_red.setOptions(op);
//     problem_red.init(H_red, g_red, A_red, NULL, NULL, lb_red, ub_red, nWSR=100);
//     problem_red.getPrimalSolution(q_red);
#include <qpOASES/include/qpOASES.hpp>

extern "C" {

// returns the init() return value; *rc_primal gets getPrimalSolution()'s return
// value, *nwsr is in/out exactly like qpOASES' nWSR argument.
int qpoases_ref_solve(int nv, int nc, const double* H, const double* g,
                      const double* A, const double* lb, const double* ub,
                      int* nwsr, double* x, int* rc_primal, double* objective)
{
  qpOASES::QProblem problem_red(nv, nc);
  qpOASES::Options op;
  op.setToMPC();
  op.printLevel = qpOASES::PL_NONE;
  problem_red.setOptions(op);
  qpOASES::int_t nWSR = *nwsr;
  int rval = problem_red.init(H, g, A, NULL, NULL, lb, ub, nWSR);
  int rval2 = problem_red.getPrimalSolution(x);
  *nwsr = (int)nWSR;
  if (rc_primal) *rc_primal = rval2;
  if (objective) *objective = problem_red.getObjVal();
  return rval;
}

int qpoases_ref_successful_return(void) { return (int)qpOASES::SUCCESSFUL_RETURN; }
int qpoases_ref_max_nwsr_return(void) { return (int)qpOASES::RET_MAX_NWSR_REACHED; }

}  // extern "C"
