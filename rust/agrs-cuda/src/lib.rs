//! Device storage for autograd.rs: what `StorageType::CudaData(CudaData<f32>)` (reference src/tensor/storage.rs:5-45)
//! becomes once the stub `CudaData { ptr }` gets a real allocation, host-side shape metadata and a `Drop`.
//! NOT compiled in the build image (no Rust toolchain); written against rust/agrs-sys, which is generated from the same
//! headers the C++ engine is built and tested against.  Every method below names the reference arm it fills.
use agrs_sys as sys;
use std::ffi::CStr;
use std::os::raw::c_void;
use std::rc::Rc;

/// One device context (stream + memory pool).  `Rc`: the reference's tensors are single-threaded (`Rc<RefCell<..>>`).
pub struct Ctx {
    raw: *mut sys::agrs_ctx,
}

impl Ctx {
    pub fn new(device: i32) -> Rc<Ctx> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { sys::agrs_ctx_create(device, &mut raw) };
        assert_eq!(rc, sys::AGRS_OK, "no sm_100 CUDA device available (there is no CPU fallback)");
        Rc::new(Ctx { raw })
    }
    pub fn raw(&self) -> *mut sys::agrs_ctx {
        self.raw
    }
    /// Non-zero status -> panic with the library's message: the reference's error convention is panic!/unwrap.
    pub fn check(&self, rc: i32) {
        if rc != sys::AGRS_OK {
            let msg = unsafe { CStr::from_ptr(sys::agrs_last_error(self.raw)) }.to_string_lossy().into_owned();
            panic!("{msg}");
        }
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::agrs_ctx_destroy(self.raw) };
    }
}

/// storage.rs:5-8 `CudaData<T> { ptr }`, with the metadata whose accessors are `todo!()` there (storage.rs:49-55, :81).
pub struct CudaData {
    pub ptr: *mut f32,
    pub shape: Vec<usize>,
    /// set by `t()` / `swap_axes` on a 2-D array (tensor.rs:165-186): consumed as an operand-major flag by `dot`
    pub transposed: bool,
    pub ctx: Rc<Ctx>,
}

impl CudaData {
    pub fn alloc(ctx: &Rc<Ctx>, shape: &[usize]) -> CudaData {
        let n: usize = shape.iter().product();
        let mut p: *mut c_void = std::ptr::null_mut();
        ctx.check(unsafe { sys::agrs_alloc(ctx.raw(), n * 4, &mut p) });
        CudaData { ptr: p as *mut f32, shape: shape.to_vec(), transposed: false, ctx: ctx.clone() }
    }
    /// `Tensor::from(ArrayD)` (tensor.rs:37-49)
    pub fn from_host(ctx: &Rc<Ctx>, data: &[f32], shape: &[usize]) -> CudaData {
        let d = CudaData::alloc(ctx, shape);
        ctx.check(unsafe { sys::agrs_upload(ctx.raw(), d.ptr as *mut c_void, data.as_ptr() as *const c_void, data.len() * 4) });
        ctx.check(unsafe { sys::agrs_sync(ctx.raw()) });
        d
    }
    /// `data()` / `into_raw_vec` (storage.rs:139-177): blocks
    pub fn to_host(&self) -> Vec<f32> {
        let mut v = vec![0f32; self.len()];
        self.ctx.check(unsafe { sys::agrs_download(self.ctx.raw(), v.as_mut_ptr() as *mut c_void, self.ptr as *const c_void, v.len() * 4) });
        v
    }
    pub fn len(&self) -> usize {
        self.shape.iter().product()
    }
    pub fn ndim(&self) -> usize {
        self.shape.len()
    }
    /// storage.rs:75
    pub fn fill(&mut self, v: f32) {
        self.ctx.check(unsafe { sys::agrs_fill_f32(self.ctx.raw(), self.ptr, v, self.len()) });
    }
    /// the CUDA arm of `Tensor::dot` (tensor.rs:319): transposes are operand majors, never copies
    pub fn dot(&self, other: &CudaData, bias: Option<&CudaData>) -> CudaData {
        assert!(self.ndim() <= 2 && other.ndim() <= 2, "Ndarray only supports 1d/2d matmul!");
        let (m, k, n) = (self.shape[0] as i64, self.shape[1] as i64, other.shape[1] as i64);
        assert_eq!(k, other.shape[0] as i64, "inputs are not compatible for matrix multiplication");
        let out = CudaData::alloc(&self.ctx, &[m as usize, n as usize]);
        let (ta, tb) = (self.transposed as i32, other.transposed as i32);
        let rc = unsafe {
            sys::agrs_gemm_f32(self.ctx.raw(), ta, tb, m, n, k, self.ptr, if self.transposed { m } else { k }, other.ptr,
                               if other.transposed { k } else { n }, out.ptr, n, bias.map_or(std::ptr::null(), |b| b.ptr as *const f32))
        };
        self.ctx.check(rc);
        out
    }
    /// `dot` + bias followed by ReLU (act = 1) or Sigmoid (act = 2) as one call: returns (pre-activation, activation).  What
    /// `Layer::forward_default` (layers.rs:24-43) can call when a Linear operator feeds an activation; the activation's node still
    /// gets the pre-activation as its variable (operators.rs:120-129, :214-224).
    pub fn dot_bias_act(&self, other: &CudaData, bias: &CudaData, act: i32) -> (CudaData, CudaData) {
        assert!(self.ndim() <= 2 && other.ndim() <= 2, "Ndarray only supports 1d/2d matmul!");
        let (m, k, n) = (self.shape[0] as i64, self.shape[1] as i64, other.shape[1] as i64);
        assert_eq!(k, other.shape[0] as i64, "inputs are not compatible for matrix multiplication");
        let z = CudaData::alloc(&self.ctx, &[m as usize, n as usize]);
        let y = CudaData::alloc(&self.ctx, &[m as usize, n as usize]);
        let (ta, tb) = (self.transposed as i32, other.transposed as i32);
        let rc = unsafe {
            sys::agrs_gemm_bias_act_f32(self.ctx.raw(), ta, tb, m, n, k, self.ptr, if self.transposed { m } else { k }, other.ptr,
                                        if other.transposed { k } else { n }, z.ptr, n, bias.ptr as *const f32, act, y.ptr, n)
        };
        self.ctx.check(rc);
        (z, y)
    }
    /// operators.rs:115 (ReLU forward) and :103 (backward_dispatch)
    pub fn relu(&self) -> CudaData {
        let out = CudaData::alloc(&self.ctx, &self.shape);
        self.ctx.check(unsafe { sys::agrs_relu_fwd_f32(self.ctx.raw(), self.ptr, out.ptr, self.len()) });
        out
    }
    pub fn relu_backward(&self, grad: &CudaData) -> CudaData {
        let out = CudaData::alloc(&self.ctx, &grad.shape);
        self.ctx.check(unsafe { sys::agrs_relu_bwd_f32(self.ctx.raw(), self.ptr, grad.ptr, out.ptr, grad.len()) });
        out
    }
    /// storage.rs:94-99 `sum_axis` (db = g.sum_axis(0), operators.rs:161)
    pub fn sum_axis(&self, axis: usize) -> CudaData {
        let outer: usize = self.shape[..axis].iter().product();
        let inner: usize = self.shape[axis + 1..].iter().product();
        let mut shape = self.shape.clone();
        shape.remove(axis);
        let out = CudaData::alloc(&self.ctx, &shape);
        self.ctx.check(unsafe { sys::agrs_sum_axis_f32(self.ctx.raw(), self.ptr, outer as i64, self.shape[axis] as i64, inner as i64, out.ptr) });
        out
    }
    /// `SGD::step` (optim.rs:25-34): w -= lr * g, g broadcast onto w when shorter (bias [1,O] -= [O])
    pub fn sgd_step(&mut self, grad: &CudaData, lr: f32) {
        self.ctx.check(unsafe { sys::agrs_sgd_step_f32(self.ctx.raw(), self.ptr, grad.ptr, lr, self.len(), grad.len()) });
    }
}

/// `#[derive(Clone)]` on the stub would alias the raw pointer: a clone is a deep device copy, as `ArrayD::clone`
/// (tensor.rs:148, :160).
impl Clone for CudaData {
    fn clone(&self) -> CudaData {
        // (no struct-update syntax here: moving fields out of a value whose type implements Drop is error E0509)
        let mut out = CudaData::alloc(&self.ctx, &self.shape);
        self.ctx.check(unsafe { sys::agrs_copy(self.ctx.raw(), out.ptr as *mut c_void, self.ptr as *const c_void, self.len() * 4) });
        out.transposed = self.transposed;
        out
    }
}

/// storage.rs:10-14: the empty `Drop` becomes a stream-ordered free (safe with kernels still in flight)
impl Drop for CudaData {
    fn drop(&mut self) {
        unsafe { sys::agrs_free(self.ctx.raw(), self.ptr as *mut c_void) };
    }
}
