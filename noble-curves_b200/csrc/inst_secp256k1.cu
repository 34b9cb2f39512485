// Instantiation of the MSM / scalar-multiplication engine for one curve.
#include "engine.cuh"
namespace nmsm {
NMSM_DEFINE_ENGINE(engine_secp256k1, CurveSecp256k1)
}
