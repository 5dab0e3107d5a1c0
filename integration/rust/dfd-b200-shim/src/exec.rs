//! `GpuRepartitionExec`: DataFusion's `RepartitionExec(Hash)` served by the B200 operator behind the C ABI.
//!
//! Reference behaviour preserved (DESIGN.md §1 invariants): same schema, same `Partitioning::Hash(exprs, n)` advertised,
//! rows of destination `p` are exactly the rows with `create_hashes(keys) % n == p`, in input order per input batch
//! sequence; an error reaches every output partition (worker_connection_pool.rs:393-397).

use std::any::Any;
use std::ffi::{CStr, CString};
use std::fmt::Formatter;
use std::ptr;
use std::sync::{Arc, Mutex};

use arrow::array::{Array, RecordBatch, StructArray};
use arrow::datatypes::SchemaRef;
use arrow::error::ArrowError;
use arrow::ffi::{to_ffi, FFI_ArrowSchema};
use arrow::ffi_stream::{ArrowArrayStreamReader, FFI_ArrowArrayStream};
use datafusion::error::{DataFusionError, Result};
use datafusion::execution::TaskContext;
use datafusion::physical_expr::expressions::Column;
use datafusion::physical_plan::repartition::RepartitionExec;
use datafusion::physical_plan::stream::RecordBatchStreamAdapter;
use datafusion::physical_plan::{
    DisplayAs, DisplayFormatType, ExecutionPlan, Partitioning, PlanProperties, SendableRecordBatchStream,
};
use futures::StreamExt;
use tokio_stream::wrappers::ReceiverStream;

use crate::ffi;

/// `dfd_last_error()` + status code → `DataFusionError` (INTEGRATION.md §3 "Error mapping").
pub(crate) fn status(code: i32, what: &str) -> Result<()> {
    if code == ffi::DFD_OK {
        return Ok(());
    }
    let msg = unsafe {
        let p = ffi::dfd_last_error();
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    };
    let text = format!("{what}: {msg}");
    Err(match code {
        ffi::DFD_ERR_INVALID_ARGUMENT => DataFusionError::Plan(text),
        ffi::DFD_ERR_UNSUPPORTED => DataFusionError::NotImplemented(text),
        ffi::DFD_ERR_OOM | ffi::DFD_ERR_CAPACITY => DataFusionError::ResourcesExhausted(text),
        ffi::DFD_ERR_CUDA | ffi::DFD_ERR_NCCL => DataFusionError::Execution(text),
        _ => DataFusionError::Internal(text),
    })
}

/// One `dfd_ctx` per GPU per worker process (reference `Worker`, src/worker/worker_service.rs:39-49). Every entry point of
/// the library is callable from any thread; calls on one context are serialised inside.
pub struct GpuContext {
    raw: *mut ffi::dfd_ctx,
}
unsafe impl Send for GpuContext {}
unsafe impl Sync for GpuContext {}

impl GpuContext {
    pub fn try_new(device: i32) -> Result<Arc<Self>> {
        let mut raw = ptr::null_mut();
        status(unsafe { ffi::dfd_ctx_create(device, &mut raw) }, "dfd_ctx_create")?;
        Ok(Arc::new(Self { raw }))
    }
}
impl Drop for GpuContext {
    fn drop(&mut self) {
        unsafe { ffi::dfd_ctx_destroy(self.raw) }
    }
}

#[derive(Clone, Copy, Debug, Default)]
pub struct GpuRepartitionOptions {
    /// rows per device chunk (0 = the library default, 4 Mi); input batches of any size are coalesced into chunks
    pub chunk_rows: i64,
    /// > 0: hard bound on pinned output chunks — `push` blocks until a consumer releases one (back-pressure,
    /// the role of the byte budget at worker_connection_pool.rs:151-153)
    pub max_pinned_chunks: i32,
}

/// The operator handle, shared by the feeder and every output stream; destroyed with its last user.
struct Operator {
    raw: *mut ffi::dfd_repartition_exec,
    _ctx: Arc<GpuContext>, // the context must outlive the operator
}
unsafe impl Send for Operator {}
unsafe impl Sync for Operator {}
impl Drop for Operator {
    fn drop(&mut self) {
        unsafe { ffi::dfd_repartition_exec_destroy(self.raw) }
    }
}

#[derive(Debug)]
pub struct GpuRepartitionExec {
    input: Arc<dyn ExecutionPlan>,
    key_columns: Vec<i32>,
    num_partitions: usize,
    properties: Arc<PlanProperties>, // cloned from the RepartitionExec being replaced: same partitioning, same eq-properties
    ctx: Arc<GpuContext>,
    options: GpuRepartitionOptions,
    /// created by the first `execute` call, shared by all output partitions (like RepartitionExec's lazily started state)
    state: Arc<Mutex<Option<Arc<Operator>>>>,
}

impl std::fmt::Debug for GpuContext {
    fn fmt(&self, f: &mut Formatter<'_>) -> std::fmt::Result {
        write!(f, "GpuContext({:p})", self.raw)
    }
}
impl std::fmt::Debug for Operator {
    fn fmt(&self, f: &mut Formatter<'_>) -> std::fmt::Result {
        write!(f, "dfd_repartition_exec({:p})", self.raw)
    }
}

/// `Hash(exprs, n)` whose expressions are all plain `Column`s → their indices; anything else is left to the CPU operator.
pub fn hash_key_columns(partitioning: &Partitioning) -> Option<(Vec<i32>, usize)> {
    match partitioning {
        Partitioning::Hash(exprs, n) => {
            let mut cols = Vec::with_capacity(exprs.len());
            for e in exprs {
                cols.push(e.as_any().downcast_ref::<Column>()?.index() as i32);
            }
            Some((cols, *n))
        }
        _ => None,
    }
}

/// Pure host check (no GPU call): `dfd_repartition_supported` — every column can be moved and every KEY column hashed the
/// way DataFusion hashes it.
pub fn repartition_supported(schema: &SchemaRef, key_columns: &[i32]) -> bool {
    match FFI_ArrowSchema::try_from(schema.as_ref()) {
        Ok(s) => unsafe { ffi::dfd_repartition_supported(&s, key_columns.as_ptr(), key_columns.len() as i32) == ffi::DFD_OK },
        Err(_) => false,
    }
}

impl GpuRepartitionExec {
    /// `None` when this `RepartitionExec` is not one the GPU operator serves (not hash, computed key expressions,
    /// unsupported column types): the hook then leaves the node alone.
    pub fn try_from_repartition(r: &RepartitionExec, ctx: Arc<GpuContext>, options: GpuRepartitionOptions) -> Option<Self> {
        let (key_columns, num_partitions) = hash_key_columns(r.partitioning())?;
        if !repartition_supported(&r.schema(), &key_columns) {
            return None;
        }
        Some(Self {
            input: Arc::clone(r.input()),
            key_columns,
            num_partitions,
            properties: r.properties().clone(),
            ctx,
            options,
            state: Arc::new(Mutex::new(None)),
        })
    }

    pub fn key_columns(&self) -> &[i32] {
        &self.key_columns
    }
    pub fn num_partitions(&self) -> usize {
        self.num_partitions
    }
    pub fn options(&self) -> GpuRepartitionOptions {
        self.options
    }

    /// First caller creates the operator and starts ONE feeder: every input partition is polled concurrently (like
    /// RepartitionExec's per-input tasks) and funnelled through a bounded channel into a single blocking thread, because
    /// `push` / `finish` are single-producer and may block on the device pipeline.
    fn operator(&self, context: &Arc<TaskContext>) -> Result<Arc<Operator>> {
        let mut guard = self.state.lock().unwrap();
        if let Some(op) = guard.as_ref() {
            return Ok(Arc::clone(op));
        }
        let schema = FFI_ArrowSchema::try_from(self.input.schema().as_ref()).map_err(DataFusionError::from)?;
        let opts = ffi::dfd_exec_options {
            chunk_rows: self.options.chunk_rows,
            max_pinned_chunks: self.options.max_pinned_chunks,
            ..Default::default()
        };
        let mut raw = ptr::null_mut();
        status(
            unsafe {
                ffi::dfd_repartition_exec_create(
                    self.ctx.raw,
                    &schema,
                    self.key_columns.as_ptr(),
                    self.key_columns.len() as i32,
                    self.num_partitions as u32,
                    &opts,
                    &mut raw,
                )
            },
            "dfd_repartition_exec_create",
        )?;
        let op = Arc::new(Operator { raw, _ctx: Arc::clone(&self.ctx) });

        let n_inputs = self.input.output_partitioning().partition_count();
        let mut inputs = Vec::with_capacity(n_inputs);
        for p in 0..n_inputs {
            inputs.push(self.input.execute(p, Arc::clone(context))?);
        }
        let (tx, mut rx) = tokio::sync::mpsc::channel::<Result<RecordBatch>>(2 * n_inputs.max(1));
        tokio::spawn(async move {
            let mut merged = futures::stream::select_all(inputs);
            while let Some(item) = merged.next().await {
                let failed = item.is_err();
                if tx.send(item).await.is_err() || failed {
                    break;
                }
            }
        });
        let feeder = Arc::clone(&op);
        tokio::task::spawn_blocking(move || {
            // an input error (or a failed export) fails EVERY output partition with the input's message, as
            // RepartitionExec does for its output channels; a failed push has already done so inside the library
            let abort = |why: String| {
                let msg = CString::new(why.replace('\0', " ")).unwrap_or_default();
                unsafe { ffi::dfd_repartition_exec_abort(feeder.raw, msg.as_ptr()) };
            };
            while let Some(item) = rx.blocking_recv() {
                let batch = match item {
                    Ok(b) => b,
                    Err(e) => return abort(e.to_string()),
                };
                let data = StructArray::from(batch).into_data();
                let (mut array, _schema) = match to_ffi(&data) {
                    Ok(x) => x,
                    Err(e) => return abort(e.to_string()),
                };
                // ownership of the exported array moves to the operator (it clears `release` in our copy)
                if unsafe { ffi::dfd_repartition_exec_push(feeder.raw, &mut array) } != ffi::DFD_OK {
                    return;
                }
            }
            unsafe { ffi::dfd_repartition_exec_finish(feeder.raw) };
        });
        *guard = Some(Arc::clone(&op));
        Ok(op)
    }
}

impl DisplayAs for GpuRepartitionExec {
    fn fmt_as(&self, _t: DisplayFormatType, f: &mut Formatter) -> std::fmt::Result {
        write!(
            f,
            "GpuRepartitionExec: partitioning=Hash({:?}, {}), input_partitions={}",
            self.key_columns,
            self.num_partitions,
            self.input.output_partitioning().partition_count()
        )
    }
}

impl ExecutionPlan for GpuRepartitionExec {
    fn name(&self) -> &str {
        "GpuRepartitionExec"
    }

    fn as_any(&self) -> &dyn Any {
        self
    }

    fn properties(&self) -> &Arc<PlanProperties> {
        &self.properties
    }

    fn children(&self) -> Vec<&Arc<dyn ExecutionPlan>> {
        vec![&self.input]
    }

    fn with_new_children(self: Arc<Self>, mut children: Vec<Arc<dyn ExecutionPlan>>) -> Result<Arc<dyn ExecutionPlan>> {
        if children.len() != 1 {
            return Err(DataFusionError::Internal("GpuRepartitionExec takes exactly one child".into()));
        }
        Ok(Arc::new(Self {
            input: children.remove(0),
            key_columns: self.key_columns.clone(),
            num_partitions: self.num_partitions,
            properties: Arc::clone(&self.properties),
            ctx: Arc::clone(&self.ctx),
            options: self.options,
            state: Arc::new(Mutex::new(None)),
        }))
    }

    fn execute(&self, partition: usize, context: Arc<TaskContext>) -> Result<SendableRecordBatchStream> {
        if partition >= self.num_partitions {
            return Err(DataFusionError::Internal(format!("partition {partition} out of range [0, {})", self.num_partitions)));
        }
        let op = self.operator(&context)?;
        let mut c_stream = FFI_ArrowArrayStream::empty();
        status(
            unsafe { ffi::dfd_repartition_exec_execute(op.raw, partition as u32, &mut c_stream) },
            "dfd_repartition_exec_execute",
        )?;
        let mut reader = ArrowArrayStreamReader::try_new(c_stream).map_err(DataFusionError::from)?;
        // get_next BLOCKS until this destination has a batch (or the input is finished): poll it on a blocking thread
        let (tx, rx) = tokio::sync::mpsc::channel::<Result<RecordBatch>>(2);
        tokio::task::spawn_blocking(move || {
            let _keep_operator_alive = op;
            for item in &mut reader {
                let item = item.map_err(|e: ArrowError| DataFusionError::from(e));
                let failed = item.is_err();
                if tx.blocking_send(item).is_err() || failed {
                    break; // consumer dropped the stream, or the operator failed (EIO from get_next)
                }
            }
        });
        Ok(Box::pin(RecordBatchStreamAdapter::new(self.schema(), ReceiverStream::new(rx))))
    }
}
