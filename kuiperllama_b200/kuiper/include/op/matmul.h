// Kept for include-path compatibility with the reference (kuiper/include/op/matmul.h): the operator
// classes of the decode path are declared together in op/decoder_layers.h.
#pragma once
#include "decoder_layers.h"
