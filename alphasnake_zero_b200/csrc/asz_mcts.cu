// placeholder until the search subsystem lands
#include "asz_engine.hpp"
namespace asz {
int search_create(asz_engine*) { return ASZ_OK; }
void search_destroy(asz_engine*) {}
}  // namespace asz
