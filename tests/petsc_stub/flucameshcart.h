#include "petsc_fluca_stub.h"
