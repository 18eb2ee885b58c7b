// Forwarding header: lets sources written against the reference's src/sph_engine.h (and through it
// src/particle.h, src/kernels.h, src/spatial_hash.h) compile against the B200 host shell unchanged.
#pragma once
#include "sph_host.h"
