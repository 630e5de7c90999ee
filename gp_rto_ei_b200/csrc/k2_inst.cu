// Translation unit holding the K2 register-tile factor kernel instances (see gprto_launch.h).
#include "k2_factor_la.cuh"
