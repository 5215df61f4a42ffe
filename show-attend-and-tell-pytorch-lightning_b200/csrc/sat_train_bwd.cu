#include "sat_common.cuh"
extern "C" int sat_train_backward(const SatDims* d, const SatWeights* w, SatTrainBuffers* b, void* stream) {
  (void)d; (void)w; (void)b; (void)stream;
  SAT_REQUIRE(false, "sat_train_backward: not built yet");
}
