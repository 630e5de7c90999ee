// Translation unit holding every K3 register-resident kernel instance (see gprto_launch.h).
#define GPRTO_K3_TU 1
#include "k3_nlml_reg.cuh"
#include "k3_nlml_la.cuh"
