#include "../../mini_ptree.h"
