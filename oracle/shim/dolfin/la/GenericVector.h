#include "dolfin_stub.h"
