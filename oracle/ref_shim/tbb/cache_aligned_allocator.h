// Forwarding header of the serial oneTBB stand-in (see shim_core.h). Test infrastructure only.
#pragma once
#include "shim_core.h"
