// rs_static_inst.cu -- one instantiation of the static-topology step kernel.  Compiled once per
// (RS_TOPO, RS_TOPO_ID, RS_MC, RS_PROF, RS_WPC) by the Makefile so the big kernels build in parallel; each object
// registers its launcher with rs_step.cu at load time.
#include "rs_kernels.cuh"
#ifndef RS_TOPO
#error "compile with -DRS_TOPO=TopoX -DRS_TOPO_ID=n -DRS_MC=.. -DRS_PROF=0|1 -DRS_WPC=k"
#endif
static int launch(const StaticLaunch& a) { return launch_static<RS_TOPO, RS_MC, RS_PROF != 0, RS_WPC>(a); }
namespace { struct Reg { Reg() { rs_register_static(RS_TOPO_ID, RS_PROF, RS_WPC, launch); } } reg; }
// physics-only sub-steps for the cube topology (rs_world_substeps of a world with FlagrunHarder's cube), once per topology
#if RS_TOPO_ID == 6 && RS_PROF == 0 && RS_WPC == 1
static int launch_sub(const SubstepsLaunch& a) { return launch_substeps_static<RS_TOPO, RS_MC>(a); }
namespace { struct RegSub { RegSub() { rs_register_substeps(RS_TOPO_ID, launch_sub); } } reg_sub; }
#endif
