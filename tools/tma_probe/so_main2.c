#include <dlfcn.h>
#include <stdio.h>
int main(int argc, char** argv) {
  void* h = dlopen(argv[1], RTLD_NOW);
  if (!h) { printf("dlopen: %s\n", dlerror()); return 1; }
  int (*f)(float*, int, int, float*, void*) = (int (*)(float*, int, int, float*, void*))dlsym(h, "osr_debug_tma_min");
  int r = f(0, 120, 1280, 0, (void*)1);
  printf("single-runtime min rc=%d\n", r);
  return 0;
}
