// sat_internal.h — the few hooks sat_train.cu needs into the (private) handle of sat_api.cu.
#pragma once
#include "../../include/sat_b200.h"

int sat_fail(int code, const char* fmt, ...);
const sat_dims* sat_handle_dims(sat_handle* h);
void** sat_handle_train_slot(sat_handle* h);
void sat_handle_set_train_free(sat_handle* h, void (*fn)(void*));
