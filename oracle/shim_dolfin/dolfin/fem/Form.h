#include "dolfin_mini.h"
