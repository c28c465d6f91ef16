// inst_1d_a.cu -- explicit kernel instantiations for a group of 1-D functors (split so the
// translation units compile in parallel).
#include "kernels_1d.cuh"
namespace gkq {
#define GKQ_PICK(ID, NAME, NP, FL) GKQ_PICK_##ID(ID, NAME, NP, FL)
#define GKQ_PICK_F1_SQUARE(ID, NAME, NP, FL) GKQ_INSTANTIATE_F1(ID, NAME, NP, FL)
#define GKQ_PICK_F1_RECIP_P(ID, NAME, NP, FL) GKQ_INSTANTIATE_F1(ID, NAME, NP, FL)
#define GKQ_PICK_F1_LORENTZ_P(ID, NAME, NP, FL) GKQ_INSTANTIATE_F1(ID, NAME, NP, FL)
#define GKQ_PICK_F1_EXPCOS_P(ID, NAME, NP, FL)
#define GKQ_PICK_F1_GAUSS_P(ID, NAME, NP, FL)
#define GKQ_PICK_F1_ABSPOW_P(ID, NAME, NP, FL)
#define GKQ_PICK_F1_LOGABS_P(ID, NAME, NP, FL)
#define GKQ_PICK_F1_F1(ID, NAME, NP, FL)
#define GKQ_PICK_F1_F2(ID, NAME, NP, FL)
#define GKQ_PICK_F1_F3(ID, NAME, NP, FL)
#define GKQ_PICK_F1_F4(ID, NAME, NP, FL)
#define GKQ_PICK_F1_F5(ID, NAME, NP, FL)
#define GKQ_PICK_F1_F6(ID, NAME, NP, FL) GKQ_INSTANTIATE_F1(ID, NAME, NP, FL)
#define GKQ_PICK_F1_EXP_P(ID, NAME, NP, FL)
#define GKQ_PICK_F1_SQRT_SHIFT_P(ID, NAME, NP, FL) GKQ_INSTANTIATE_F1(ID, NAME, NP, FL)
#define GKQ_PICK_F1_INV_ABS_SHIFT_P(ID, NAME, NP, FL) GKQ_INSTANTIATE_F1(ID, NAME, NP, FL)
#define GKQ_PICK_F1_COS_P(ID, NAME, NP, FL)
#define GKQ_PICK_F1_SIN_P(ID, NAME, NP, FL)
#define GKQ_PICK_F1_SEMICIRCLE_P(ID, NAME, NP, FL)
#define GKQ_PICK_F1_SINPOW_P(ID, NAME, NP, FL)
GKQ_F1_LIST(GKQ_PICK)
}  // namespace gkq
