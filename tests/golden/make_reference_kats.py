"""Writes tests/golden/reference_kats.json: the golden vectors held by the reference's OWN
unit tests for the hot path, transcribed literal-for-literal (values only, no code) with the
file:line they come from.  /root/reference is not needed to run this script; it exists so
the fixture's provenance is reviewable.  All vectors are small integers => exact under any
summation order, which is why they can pin both the oracle and the CUDA path bit-for-bit.
"""
import json, os

K = {}

# operators.rs:486-506  test_relu_on_mat: a=[[-1,-2],[3,4]]; a+=a; a.flat[0]=1; ReLU
K["relu_on_mat"] = {"ref": "src/operators/operators.rs:486-506",
                    "x": [[1., -4.], [6., 8.]], "y": [[1., 0.], [6., 8.]]}

# operators.rs:508-521  test_linear
K["linear"] = {"ref": "src/operators/operators.rs:508-521",
               "x": [[1., 1.]], "w": [[1., 1.], [1., 1.]], "b": [[1., 1.]], "y": [[3., 3.]]}

# operators.rs:523-545  test_im2col (no pad): im = range(1,49).reshape(1,3,4,4), k=2,s=1,p=0.
# The test builds `expected` as a 12x9 matrix M and compares with M.t().into_shape((1,9,12))
# on an F-contiguous view, i.e. col[0,p,q] == M[q][p].
M1 = [
    [1., 2., 3., 5., 6., 7., 9., 10., 11.],
    [2., 3., 4., 6., 7., 8., 10., 11., 12.],
    [5., 6., 7., 9., 10., 11., 13., 14., 15.],
    [6., 7., 8., 10., 11., 12., 14., 15., 16.],
    [17., 18., 19., 21., 22., 23., 25., 26., 27.],
    [18., 19., 20., 22., 23., 24., 26., 27., 28.],
    [21., 22., 23., 25., 26., 27., 29., 30., 31.],
    [22., 23., 24., 26., 27., 28., 30., 31., 32.],
    [33., 34., 35., 37., 38., 39., 41., 42., 43.],
    [34., 35., 36., 38., 39., 40., 42., 43., 44.],
    [37., 38., 39., 41., 42., 43., 45., 46., 47.],
    [38., 39., 40., 42., 43., 44., 46., 47., 48.],
]
K["im2col_nopad"] = {"ref": "src/operators/operators.rs:523-545",
                     "im_shape": [1, 3, 4, 4], "im_range": [1, 49], "k": 2, "stride": 1, "pad": 0,
                     "col_shape": [1, 9, 12],
                     "col": [[M1[q][p] for q in range(12)] for p in range(9)]}

# operators.rs:546-565  test_im2col (pad=1): im = range(1,9).reshape(1,2,2,2), k=2,s=1,p=1
K["im2col_pad"] = {"ref": "src/operators/operators.rs:546-565",
                   "im_shape": [1, 2, 2, 2], "im_range": [1, 9], "k": 2, "stride": 1, "pad": 1,
                   "col_shape": [1, 9, 8],
                   "col": [
                       [0., 0., 0., 1., 0., 0., 0., 5.],
                       [0., 0., 1., 2., 0., 0., 5., 6.],
                       [0., 0., 2., 0., 0., 0., 6., 0.],
                       [0., 1., 0., 3., 0., 5., 0., 7.],
                       [1., 2., 3., 4., 5., 6., 7., 8.],
                       [2., 0., 4., 0., 6., 0., 8., 0.],
                       [0., 3., 0., 0., 0., 7., 0., 0.],
                       [3., 4., 0., 0., 7., 8., 0., 0.],
                       [4., 0., 0., 0., 8., 0., 0., 0.]]}

# operators.rs:568-602  test_conv2d: ones everywhere, B=1,Ci=2,Co=3,k=2 -> all outputs 9.0;
# backward: only shapes are asserted (grads[i].shape == xs[i].shape for i in 0..3)
K["conv2d_ones"] = {"ref": "src/operators/operators.rs:568-602",
                    "im_shape": [1, 2, 4, 4], "filt_shape": [2, 2, 3], "bias_shape": [3],
                    "c_in": 2, "c_out": 3, "k": 2, "stride": 1, "pad": 0,
                    "out_shape": [1, 3, 3, 3], "out_value": 9.0,
                    "bwd_shapes": [[1, 2, 4, 4], [2, 2, 3], [3]]}

# autograd.rs:94-119  test_single_node_graph (pins relu'(0) = 0)
K["single_node_graph"] = {"ref": "src/autograd/autograd.rs:94-119",
                          "x": [[0., -1.], [2., 3.]], "y": [[0., 0.], [2., 3.]],
                          "x_grad": [[0., 0.], [1., 1.]]}

# autograd.rs:121-171  test_simple_graph: ReLU -> Linear, backward from the [1,2] output
K["simple_graph"] = {"ref": "src/autograd/autograd.rs:121-171",
                     "x": [[0., 1.]], "w": [[1., 1.], [1., 1.]], "b": [[1., 1.]],
                     "y": [[2., 2.]], "x_grad": [[0., 2.]],
                     "w_grad": [[0., 0.], [1., 1.]], "b_grad": [1., 1.]}

# tensor_arithmetic.rs:348-399: in-place vs new-tensor arithmetic semantics
K["arith_adds"] = {"ref": "src/tensor/tensor_arithmetic.rs:348-365",
                   "note": "a=[[1,2],[3,4]] (zeros+...), see test for the exact handle dance",
                   "a": [[1., 2.], [3., 4.]], "b": [[1., 1.], [1., 1.]],
                   "a_plus_b": [[2., 3.], [4., 5.]]}
K["arith_scalar"] = {"ref": "src/tensor/tensor_arithmetic.rs:385-399",
                     "a": [[1., 2.], [3., 4.]],
                     "mul2": [[2., 4.], [6., 8.]], "div2": [[0.5, 1.], [1.5, 2.]],
                     "mul_assign3": [[3., 6.], [9., 12.]], "three_minus_a": [[2., 1.], [0., -1.]]}

# model.rs:31-62 test_mlp: parameter shapes of Linear(20,10) -> ReLU -> Linear(10,1)
K["mlp_param_shapes"] = {"ref": "src/nn/model.rs:58-61",
                         "shapes": [[20, 10], [1, 10], [10, 1], [1, 1]], "out_shape": [1, 1]}

# tensor.rs:464-502: metadata / aliasing
K["basic_from"] = {"ref": "src/tensor/tensor.rs:464-472", "shape": [2, 3], "ndim": 2, "len": 6}
K["reshape_alias"] = {"ref": "src/tensor/tensor.rs:493-502", "from": [2, 2], "to": [4]}

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
with open(out, "w") as f:
    json.dump(K, f, indent=1)
print("wrote", out, len(K), "KATs")
