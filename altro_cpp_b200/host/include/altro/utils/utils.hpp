// altro/utils/utils.hpp (B200 host mirror)
#pragma once

#include "altro/utils/assert.hpp"

#define ALTRO_UNUSED(var) (void)(var)
