// Zero-filling operator new[] for the A/B build of the reference's tvl1occflow CLI (cli/Makefile,
// target tvl1occflow_ref): Solver_wrt_chi reads eta1 / eta2 from freshly new[]-ed memory without
// initialising it (src/tvl1occflow_solvers.cpp:241-264, the file's own #warning), so the unmodified
// program's output depends on what the allocator hands back.  With this object linked in, it computes
// what a fresh process computes on large images (zero pages) at every size.  Not part of the product.
#include <cstdlib>
#include <new>

void *operator new[](std::size_t n)
{
    void *p = std::calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete[](void *p) noexcept { std::free(p); }
void operator delete[](void *p, std::size_t) noexcept { std::free(p); }
