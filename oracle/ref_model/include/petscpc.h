/* oracle/ref_model: TEST INFRASTRUCTURE.  Stands in for the PETSc header of this name; everything is in petsc_model.h */
#pragma once
#include <petsc_model.h>
