#include "petsc_shim.h"
extern int PETSC_COMM_WORLD;
