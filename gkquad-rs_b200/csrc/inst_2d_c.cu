// inst_2d_c.cu -- explicit kernel instantiations for a group of 2-D functors.
#include "kernels_2d.cuh"
namespace gkq {
#define GKQ_PICK(ID, NAME, NP, FL) GKQ_PICK_##ID(ID, NAME, NP, FL)
#define GKQ_PICK_F2_XY(ID, NAME, NP, FL)
#define GKQ_PICK_F2_RECIP_XY_P(ID, NAME, NP, FL)
#define GKQ_PICK_F2_RSQRT3(ID, NAME, NP, FL)
#define GKQ_PICK_F2_G1(ID, NAME, NP, FL)
#define GKQ_PICK_F2_G2(ID, NAME, NP, FL)
#define GKQ_PICK_F2_G3(ID, NAME, NP, FL)
#define GKQ_PICK_F2_G4(ID, NAME, NP, FL)
#define GKQ_PICK_F2_GP1(ID, NAME, NP, FL)
#define GKQ_PICK_F2_GP2(ID, NAME, NP, FL) GKQ_INSTANTIATE_F2(ID, NAME, NP, FL)
#define GKQ_PICK_F2_GP3(ID, NAME, NP, FL) GKQ_INSTANTIATE_F2(ID, NAME, NP, FL)
#define GKQ_PICK_F2_D1_P(ID, NAME, NP, FL) GKQ_INSTANTIATE_F2(ID, NAME, NP, FL)
#define GKQ_PICK_F2_HEMISPHERE_P(ID, NAME, NP, FL)
GKQ_F2_LIST(GKQ_PICK)
}  // namespace gkq
