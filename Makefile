# Builds libbgn_b200.so (sm_100a only) and the CPU-side test / oracle helpers.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
# limb counts to instantiate; `make LIMBS=17` is a quick single-size developer build
LIMBS     ?= 3 5 9 17 33
HAVE      := $(foreach l,$(LIMBS),-DBGN_HAVE_L$(l))
SRC       := bgn_b200/csrc
OBJ       := build
HDRS      := $(wildcard $(SRC)/*.cuh $(SRC)/*.h) include/bgn_b200.h
OBJS      := $(OBJ)/api.o $(foreach l,$(LIMBS),$(OBJ)/inst_a_$(l).o $(OBJ)/inst_b_$(l).o $(OBJ)/inst_c_$(l).o $(OBJ)/inst_d_$(l).o $(OBJ)/inst_e_$(l).o)
LIB       := bgn_b200/libbgn_b200.so

all: $(LIB)

$(OBJ)/inst_a_%.o: $(SRC)/inst_a.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DBGN_L=$* -c $< -o $@ 2> $(OBJ)/inst_a_$*.log || (tail -30 $(OBJ)/inst_a_$*.log; false)

$(OBJ)/inst_b_%.o: $(SRC)/inst_b.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DBGN_L=$* -c $< -o $@ 2> $(OBJ)/inst_b_$*.log || (tail -30 $(OBJ)/inst_b_$*.log; false)

$(OBJ)/inst_c_%.o: $(SRC)/inst_c.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DBGN_L=$* -c $< -o $@ 2> $(OBJ)/inst_c_$*.log || (tail -30 $(OBJ)/inst_c_$*.log; false)

$(OBJ)/inst_d_%.o: $(SRC)/inst_d.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DBGN_L=$* -c $< -o $@ 2> $(OBJ)/inst_d_$*.log || (tail -30 $(OBJ)/inst_d_$*.log; false)

$(OBJ)/inst_e_%.o: $(SRC)/inst_e.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DBGN_L=$* -c $< -o $@ 2> $(OBJ)/inst_e_$*.log || (tail -30 $(OBJ)/inst_e_$*.log; false)

# api.o depends on WHICH limb counts are linked in: re-made whenever LIMBS changes
$(OBJ)/limbs.stamp: FORCE
	@mkdir -p $(OBJ)
	@echo '$(LIMBS)' | cmp -s - $@ || echo '$(LIMBS)' > $@
FORCE:

$(OBJ)/api.o: $(SRC)/api.cu $(HDRS) $(OBJ)/limbs.stamp
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(HAVE) -c $< -o $@ 2> $(OBJ)/api.log || (tail -30 $(OBJ)/api.log; false)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

clean:
	rm -rf $(OBJ) $(LIB)

.PHONY: all clean FORCE
