#include <R.h>
