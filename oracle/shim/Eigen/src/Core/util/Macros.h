#include "../../../../mini_eigen.h"
