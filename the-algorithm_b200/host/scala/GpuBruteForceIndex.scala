package com.twitter.ann.brute_force

// Drop-in for com.twitter.ann.brute_force.BruteForceIndex (BruteForceIndex.scala:26-92) backed by the B200 engine.
// Same traits, same factory shape; swap the constructor call at the call sites listed in INTEGRATION.md.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no JVM / scalac / jni.h (SURVEY.md F4).  It is the thin,
// mechanical binding of include/b200ann.h that a maintainer adds next to the reference class; every behaviour it relies
// on is exercised through the same C ABI by tests/ (Python ctypes) and tests/cpp_host_check.cpp (C++).

import com.twitter.ann.common.Appendable
import com.twitter.ann.common.Distance
import com.twitter.ann.common.EmbeddingType._
import com.twitter.ann.common.EntityEmbedding
import com.twitter.ann.common.Metric
import com.twitter.ann.common.NeighborWithDistance
import com.twitter.ann.common.Queryable
import com.twitter.ann.common.Updatable
import com.twitter.ann.common.{Cosine, InnerProduct, L2}
import com.twitter.util.Future
import com.twitter.util.FuturePool
import java.nio.{ByteBuffer, ByteOrder}

/** JNI surface: one static native per C-ABI entry point (include/b200ann.h).  Buffers are direct ByteBuffers so the
 * GC can neither move nor free them during a call (the hazard noted at faiss/QueryableIndexAdapter.scala:128-132). */
object B200AnnNative {
  // natives are unpacked from the jar and System.load'ed exactly like swigfaiss (faiss/NativeUtils.java:93-111)
  System.loadLibrary("b200ann_jni")
  @native def create(metric: Int, dim: Int, capacityHint: Long, device: Int, flags: Int): Long // 0 => see lastError
  @native def destroy(handle: Long): Unit
  @native def appendBatch(handle: Long, ids: ByteBuffer, rows: ByteBuffer, n: Long): Int
  @native def size(handle: Long): Long
  @native def queryBatch(handle: Long, queries: ByteBuffer, b: Int, dim: Int, k: Int,
    outIds: ByteBuffer, outDist: ByteBuffer, outCount: ByteBuffer): Int
  // KnnHelper.findNearestNeighbours (scalding/offline/KnnHelper.scala:168-215) as one native job: ann_knn_join
  @native def knnJoin(metric: Int, dim: Int, device: Int, flags: Int, corpusIds: ByteBuffer, corpusRows: ByteBuffer, n: Long,
    queries: ByteBuffer, nq: Long, k: Int, corpusTileRows: Long, queryTile: Int,
    outIds: ByteBuffer, outDist: ByteBuffer, outCount: ByteBuffer): Int
  // Metric.distance for n pairs / MetricUtil.norm (common/Metric.scala:76-86, 285-289) on the device, query-path arithmetic
  @native def distancePairs(metric: Int, flags: Int, dim: Int, a: ByteBuffer, b: ByteBuffer, n: Long, out: ByteBuffer,
    device: Int): Int
  @native def normalizeRows(dim: Int, rows: ByteBuffer, n: Long, out: ByteBuffer, device: Int): Int
  // one shard's half of ComposedQueryable's fan-out (common/ShardApi.scala:72-79) when the shards are GPUs of one box:
  // publish bounds -> [cross-GPU barrier] -> score against the global bound (ann_query_seed_device / _finish_device)
  @native def querySeedDevice(handle: Long, dQueries: Long, b: Int, dim: Int, k: Int, dSeedKeys: Long, stream: Long): Int
  @native def queryFinishDevice(handle: Long, dQueries: Long, b: Int, dim: Int, k: Int, peerSeedKeys: ByteBuffer, world: Int,
    dOutIds: Long, dOutDist: Long, dOutCount: Long, stream: Long): Int
  // one process, several GPUs: ShardedAppendable + ComposedQueryable as ONE native handle (ann_sharded_*; ShardApi.scala:34-87)
  @native def shardedCreate(metric: Int, dim: Int, capacityHint: Long, flags: Int, devices: Array[Int]): Long
  @native def shardedDestroy(handle: Long): Unit
  @native def shardedAppendBatch(handle: Long, ids: ByteBuffer, rows: ByteBuffer, n: Long): Int
  @native def shardedSize(handle: Long): Long
  @native def shardedQueryBatch(handle: Long, queries: ByteBuffer, b: Int, dim: Int, k: Int,
    outIds: ByteBuffer, outDist: ByteBuffer, outCount: ByteBuffer): Int
  // the reference's own on-disk format, read and written natively (BruteForceIndex.scala:142-161; ThriftIteratorIO.scala:14-56;
  // ShardedSerialization.scala:28-66): ann_save_directory / ann_load_directory and the sharded variants
  @native def saveDirectory(handle: Long, directory: String, idFormat: Int, layout: Int): Int
  @native def loadDirectory(metric: Int, dim: Int, device: Int, flags: Int, directory: String, idFormat: Int): Long
  @native def shardedSaveDirectory(handle: Long, directory: String, idFormat: Int, layout: Int): Int
  @native def shardedLoadDirectory(metric: Int, dim: Int, flags: Int, directory: String, idFormat: Int, devices: Array[Int]): Long
  // Updatable.update (common/Api.scala:148-150), batched: overwrite the rows stored at insertion slots (ann_update_batch)
  @native def updateBatch(handle: Long, slots: ByteBuffer, rows: ByteBuffer, n: Long): Int
  // rows [start, start + n) and their ids back into direct buffers (ann_read_rows)
  @native def readRows(handle: Long, start: Long, n: Long, outIds: ByteBuffer, outRows: ByteBuffer): Int
  // tuning / introspection by name (ann_set_option / ann_get_stat; getStat returns Long.MinValue on failure)
  @native def setOption(handle: Long, name: String, value: Long): Int
  @native def getStat(handle: Long, name: String): Long
  @native def version(): Int
  // message of the last failed call on this thread; the shim refuses direct buffers too small for a call before the
  // library could overrun them, and reports that here as well
  @native def lastError(): String
}

object GpuBruteForceIndex {
  val DataFileName = "BruteForceFileData"

  private def ordinal(metric: Metric[_]): Int = metric match { // thrift DistanceMetric, ann_common.thrift:16-19
    case L2 => 0
    case Cosine => 1
    case InnerProduct => 2
    case other => throw new IllegalArgumentException(s"metric $other is not a dense-vector metric")
  }

  /** Same shape as BruteForceIndex.apply (BruteForceIndex.scala:29-37).  `T` is carried through `idInjection`:
   * Long ids go to the device as they are (ties break by id); any other T is mapped to its insertion slot. */
  def apply[T, D <: Distance[D]](
    metric: Metric[D],
    futurePool: FuturePool,
    initialEmbeddings: Iterator[EntityEmbedding[T]] = Iterator(),
    device: Int = 0
  ): GpuBruteForceIndex[T, D] = {
    val index = new GpuBruteForceIndex[T, D](metric, futurePool, device)
    initialEmbeddings.grouped(65536).foreach(batch => index.appendBatch(batch))
    index
  }

  /** BruteForceDeserialization.fromDirectory (BruteForceDeserialization.scala:42-63) for Long ids: reads the
   * `BruteForceFileData` thrift stream written by SerializableBruteForceIndex.toDirectory -- or by `toDirectory` below --
   * straight into device memory (ann_load_directory).  The dimension is taken from the first record. */
  def fromDirectory[D <: Distance[D]](
    directory: String,
    metric: Metric[D],
    futurePool: FuturePool,
    device: Int = 0
  ): GpuBruteForceIndex[Long, D] = {
    val h = B200AnnNative.loadDirectory(ordinal(metric), 0, device, 0, directory, 0)
    if (h == 0L) throw new RuntimeException(B200AnnNative.lastError())
    val index = new GpuBruteForceIndex[Long, D](metric, futurePool, device)
    index.adopt(h)
    index
  }
}

class GpuBruteForceIndex[T, D <: Distance[D]] private (
  metric: Metric[D],
  futurePool: FuturePool,
  device: Int)
    extends Appendable[T, BruteForceRuntimeParams.type, D]
    with Queryable[T, BruteForceRuntimeParams.type, D]
    with Updatable[T]
    with AutoCloseable {

  private[this] var handle: Long = 0L
  private[this] var dim: Int = -1
  private[this] val slotTable = new java.util.ArrayList[T]() // used only when T is not Long
  private[this] var nativeIds = true
  // id -> insertion slot of its FIRST occurrence (what Updatable.update rewrites); filled as rows are appended
  private[this] val slotOf = new java.util.HashMap[T, java.lang.Long]()
  private[this] var appended: Long = 0L

  /** takes over a handle that ann_load_directory filled (Long ids); the id -> slot map is rebuilt from the device */
  private[brute_force] def adopt(loaded: Long): Unit = synchronized {
    handle = loaded
    dim = B200AnnNative.getStat(loaded, "dim").toInt
    appended = B200AnnNative.size(loaded)
    var start = 0L
    while (start < appended) {
      val n = math.min(1L << 20, appended - start)
      val ids = direct((n * 8).toInt)
      check(B200AnnNative.readRows(handle, start, n, ids, null))
      var i = 0
      while (i < n) { slotOf.putIfAbsent(ids.getLong(i * 8).asInstanceOf[T], start + i); i += 1 }
      start += n
    }
  }

  private[this] def check(rc: Int): Unit =
    if (rc != 0) throw new RuntimeException(s"b200ann error $rc: ${B200AnnNative.lastError()}")

  private[this] def ensure(d: Int): Unit = synchronized {
    if (handle == 0L) {
      handle = B200AnnNative.create(GpuBruteForceIndex.ordinal(metric), d, 0L, device, 0)
      if (handle == 0L) throw new RuntimeException(B200AnnNative.lastError())
      dim = d
    } else if (d != dim) {
      throw new IllegalArgumentException(s"embedding dimension $d != index dimension $dim")
    }
  }

  private[this] def direct(bytes: Int): ByteBuffer = ByteBuffer.allocateDirect(bytes).order(ByteOrder.nativeOrder())

  /** Batched Appendable path: one host->device copy and one kernel per batch. */
  def appendBatch(batch: Seq[EntityEmbedding[T]]): Unit = synchronized {
    if (batch.nonEmpty) {
      val d = batch.head.embedding.length
      ensure(d)
      val ids = direct(batch.size * 8)
      val rows = direct(batch.size * d * 4)
      batch.foreach { e =>
        e.id match {
          case l: Long if nativeIds => ids.putLong(l)
          case other =>
            nativeIds = false
            ids.putLong(slotTable.size.toLong)
            slotTable.add(other)
        }
        var i = 0
        while (i < d) { rows.putFloat(e.embedding(i)); i += 1 }
      }
      check(B200AnnNative.appendBatch(handle, ids, rows, batch.size.toLong))
      // only after the native call succeeded, so a failed append cannot desynchronise id and slot
      batch.foreach { e => slotOf.putIfAbsent(e.id, appended); appended += 1 }
    }
  }

  /** Updatable.update (Api.scala:148-150; hnsw/Hnsw.scala:149-182 is the in-tree implementation) for a whole batch: the
   * embedding stored under each id is overwritten in place, norms and the bf16 shadow row are refreshed (ann_update_batch).
   * An unknown id is an error, as in Hnsw.update. */
  def updateBatch(batch: Seq[EntityEmbedding[T]]): Unit = synchronized {
    if (batch.nonEmpty) {
      val slots = direct(batch.size * 8)
      val rows = direct(batch.size * dim * 4)
      batch.foreach { e =>
        val slot = slotOf.get(e.id)
        if (slot == null) throw new IllegalArgumentException(s"update of an id that was never appended: ${e.id}")
        if (e.embedding.length != dim) throw new IllegalArgumentException(s"embedding dimension ${e.embedding.length} != index dimension $dim")
        slots.putLong(slot)
        var i = 0
        while (i < dim) { rows.putFloat(e.embedding(i)); i += 1 }
      }
      check(B200AnnNative.updateBatch(handle, slots, rows, batch.size.toLong))
    }
  }

  override def update(entity: EntityEmbedding[T]): Future[Unit] = futurePool { updateBatch(Seq(entity)) }

  /** SerializableBruteForceIndex.toDirectory (BruteForceIndex.scala:142-161): `BruteForceFileData` as a TBinaryProtocol
   * PersistedEmbedding stream + `_SUCCESS`, written natively (ann_save_directory); Long ids only (8 bytes big-endian,
   * AnnInjections.scala:8-12).  BruteForceDeserialization.fromDirectory reads it back, as does `fromDirectory` above. */
  def toDirectory(directory: String): Unit = synchronized {
    if (!nativeIds) throw new UnsupportedOperationException("toDirectory needs Long ids (other id types live in the JVM-side slot table)")
    if (handle == 0L) throw new IllegalStateException("toDirectory on an index that never saw a row (dimension unknown)")
    check(B200AnnNative.saveDirectory(handle, directory, 1, 0))
  }

  // Appendable.append, BruteForceIndex.scala:48-52
  override def append(embedding: EntityEmbedding[T]): Future[Unit] = futurePool { appendBatch(Seq(embedding)) }

  override def toQueryable: Queryable[T, BruteForceRuntimeParams.type, D] = this

  /** b queries in one device call; the extra entry point callers that can batch should use (SURVEY.md 3.3). */
  def batchQueryWithDistance(
    embeddings: Seq[EmbeddingVector],
    numOfNeighbours: Int
  ): Seq[List[NeighborWithDistance[T, D]]] = {
    if (numOfNeighbours <= 0 || handle == 0L || embeddings.isEmpty) return embeddings.map(_ => Nil)
    val b = embeddings.size
    val q = direct(b * dim * 4)
    embeddings.foreach { e => var i = 0; while (i < dim) { q.putFloat(e(i)); i += 1 } }
    val outIds = direct(b * numOfNeighbours * 8)
    val outDist = direct(b * numOfNeighbours * 4)
    val outCount = direct(b * 4)
    check(B200AnnNative.queryBatch(handle, q, b, dim, numOfNeighbours, outIds, outDist, outCount))
    (0 until b).map { qi =>
      (0 until outCount.getInt(qi * 4)).map { j =>
        val raw = outIds.getLong((qi * numOfNeighbours + j) * 8)
        val id = if (nativeIds) raw.asInstanceOf[T] else slotTable.get(raw.toInt)
        NeighborWithDistance(id, metric.fromAbsoluteDistance(outDist.getFloat((qi * numOfNeighbours + j) * 4)))
      }.toList
    }
  }

  // Queryable.queryWithDistance, BruteForceIndex.scala:66-91: nearest first; ties by id instead of by heap history
  override def queryWithDistance(
    embedding: EmbeddingVector,
    numOfNeighbours: Int,
    runtimeParams: BruteForceRuntimeParams.type
  ): Future[List[NeighborWithDistance[T, D]]] =
    futurePool { batchQueryWithDistance(Seq(embedding), numOfNeighbours).head }

  // Queryable.query, BruteForceIndex.scala:56-64
  override def query(
    embedding: EmbeddingVector,
    numOfNeighbours: Int,
    runtimeParams: BruteForceRuntimeParams.type
  ): Future[List[T]] =
    queryWithDistance(embedding, numOfNeighbours, runtimeParams).map(_.map(_.neighbor))

  def size: Long = synchronized { if (handle == 0L) 0L else B200AnnNative.size(handle) }

  // explicit release, idempotent -- the swig Index.delete() contract (faiss/swig/Index.java:24-37)
  override def close(): Unit = synchronized {
    if (handle != 0L) { B200AnnNative.destroy(handle); handle = 0L }
  }
}
