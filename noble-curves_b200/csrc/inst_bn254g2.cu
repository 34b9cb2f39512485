// Instantiation of the MSM / scalar-multiplication engine for one curve.
#include "engine.cuh"
namespace nmsm {
NMSM_DEFINE_ENGINE(engine_bn254g2, CurveBn254G2)
}
