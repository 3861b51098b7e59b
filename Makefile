# Builds libsqlx.so (sm_100a only) in-tree.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
PKG       := sfmnext-impl_b200
SRC       := $(wildcard $(PKG)/csrc/*.cu)
OBJ       := $(patsubst $(PKG)/csrc/%.cu,$(PKG)/build/%.o,$(SRC))
LIB       := $(PKG)/lib/libsqlx.so

all: $(LIB)

$(PKG)/build/%.o: $(PKG)/csrc/%.cu $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/csrc/*.h) include/sqlx.h
	@mkdir -p $(PKG)/build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(PKG)/build/$*.ptxas.log || (cat $(PKG)/build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p $(PKG)/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart -lcuda

# A/B variants for kernel tuning (never loaded by the product unless SQLX_LIB_PATH points at one):
#   make variant NAME=nw3 DEFS="-DSQLX_FWD_NW=3"   ->  $(PKG)/lib/libsqlx_nw3.so      (tools/ab.py times them side by side)
variant:
	@mkdir -p $(PKG)/build_$(NAME) $(PKG)/lib
	for f in $(SRC); do o=$(PKG)/build_$(NAME)/$$(basename $$f .cu).o; \
	  $(NVCC) $(NVCCFLAGS) $(DEFS) -c $$f -o $$o 2> $$o.log || (cat $$o.log; exit 1); done
	$(NVCC) $(ARCH) -shared -o $(PKG)/lib/libsqlx_$(NAME).so $(PKG)/build_$(NAME)/*.o -lcudart -lcuda

clean:
	rm -rf $(PKG)/build $(PKG)/build_* $(PKG)/lib

.PHONY: all clean variant
