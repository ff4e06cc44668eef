// Stub of the third-party tinyexpr API (driver motion expressions are out of scope; never evaluated by the oracle).
#pragma once
struct te_expr { int dummy; };
struct te_variable { const char* name; const void* address; int type; void* context; };
inline te_expr* te_compile(const char*, const te_variable*, int, int* err) { if (err) *err = 1; return nullptr; }
inline double te_eval(const te_expr*) { return 0.0; }
inline void te_free(te_expr*) {}
