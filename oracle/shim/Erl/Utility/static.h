#include "../../mini_erl.h"
