// Instantiation of the MSM / scalar-multiplication engine for one curve.
#include "engine.cuh"
namespace nmsm {
NMSM_DEFINE_ENGINE(engine_bls381g2_any, CurveBls381G2Any)
}
