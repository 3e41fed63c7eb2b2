// Include-path compatibility: Sampler and ArgmaxSampler are declared in sampler/argmax_sampler.h.
#pragma once
#include "argmax_sampler.h"
