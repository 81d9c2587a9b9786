#include "chipmunk.h"
