#include "petsc_shim.h"
