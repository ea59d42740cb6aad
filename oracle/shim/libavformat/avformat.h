#include "../libavutil/frame.h"
