// ob_dropin.cpp — the compute entry points of the classic ODE API (dSpaceCollide,
// dCollide, dWorldQuickStep) served by the CUDA kernels through a batch of one
// world.  (filled in after the batched path; until then they report an error
// rather than compute anything on the CPU.)
#include "ob_host.h"
void ob_dropin_space_collide(dxSpace *, void *, dNearCallback *) { ob_error(0, "dSpaceCollide: drop-in path not available in this build (use dBatch*)"); }
void ob_dropin_space_collide2(dxGeom *, dxGeom *, void *, dNearCallback *) { ob_error(0, "dSpaceCollide2: drop-in path not available in this build"); }
int ob_dropin_collide(dxGeom *, dxGeom *, int, dContactGeom *, int) { ob_error(0, "dCollide: drop-in path not available in this build"); return 0; }
int ob_dropin_quickstep(dxWorld *, dReal) { ob_error(0, "dWorldQuickStep: drop-in path not available in this build (use dBatch*)"); return 0; }
