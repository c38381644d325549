// stubs.cu — temporary: entry points / executors not implemented yet.
#include "engine.h"
namespace ark {
}

#define ARK_STUB(name) { ark::set_last_error(name " is not implemented yet"); return ARK_ERR_UNSUPPORTED; }
extern "C" {
int ark_arrow_to_json_create(const char*, ark_proc_t**) ARK_STUB("ark_arrow_to_json_create")
int ark_arrow_to_json_process(ark_proc_t*, ArrowArray*, ArrowSchema*, ArrowArray*, ArrowSchema*) ARK_STUB("ark_arrow_to_json_process")
}
