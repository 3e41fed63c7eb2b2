// Include-path compatibility: allocators and Buffer are declared in base/memory.h.
#pragma once
#include "memory.h"
