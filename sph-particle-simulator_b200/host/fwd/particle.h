#pragma once
#include "sph_host.h"
