// stubs.cu — temporary: entry points / executors not implemented yet.
#include "engine.h"
namespace ark {
}

#define ARK_STUB(name) { ark::set_last_error(name " is not implemented yet"); return ARK_ERR_UNSUPPORTED; }
extern "C" {
int ark_arrow_to_json_create(const char*, ark_proc_t**) ARK_STUB("ark_arrow_to_json_create")
int ark_arrow_to_json_process(ark_proc_t*, ArrowArray*, ArrowSchema*, ArrowArray*, ArrowSchema*) ARK_STUB("ark_arrow_to_json_process")
int ark_buffer_create(const char*, const char*, const char*, ark_buf_t**) ARK_STUB("ark_buffer_create")
int ark_buffer_write(ark_buf_t*, ArrowArray*, ArrowSchema*, const char*, uint64_t) ARK_STUB("ark_buffer_write")
int ark_buffer_read(ark_buf_t*, ArrowArray*, ArrowSchema*, uint64_t*, int64_t, int64_t*) ARK_STUB("ark_buffer_read")
int ark_buffer_flush(ark_buf_t*) ARK_STUB("ark_buffer_flush")
int ark_buffer_close(ark_buf_t*) ARK_STUB("ark_buffer_close")
void ark_buffer_destroy(ark_buf_t*) {}
}
