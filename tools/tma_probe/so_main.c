#include <dlfcn.h>
#include <stdio.h>
int main(int argc, char** argv) {
  void* h = dlopen(argv[1], RTLD_NOW);
  if (!h) { printf("dlopen: %s\n", dlerror()); return 1; }
  int (*f)(int, char**) = (int (*)(int, char**))dlsym(h, "probe_main");
  return f(argc - 1, argv + 1);
}
