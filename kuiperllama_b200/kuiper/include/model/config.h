// Include-path compatibility: see model/checkpoint_file.h.
#pragma once
#include "checkpoint_file.h"
