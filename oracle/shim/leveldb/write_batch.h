// Reference-harness stand-in for LevelDB (all five headers live in db.h).
#pragma once
#include "leveldb/db.h"
