// DBImprovedGpu.cs -- drop-in for vtkPointCloud.DBImproved (BaseClass/DBImproved.cs:8-116): same public
// fields, same dbscan(List<Point3D>, double, int) signature and the same mutation contract, computed by
// libvpc on the GPU.  Swap `new DBImproved()` for `new DBImprovedGpu()` at FrmMain.cs:1508, :2785 and
// Tools.cs:591.  Source only (no .NET toolchain in the build image).
using System;
using System.Collections.Generic;

namespace vtkPointCloud
{
    public class DBImprovedGpu : IDisposable
    {
        public int clusterAmount = 0;   // DBImproved.cs:10
        public int pointsAmount = 0;    // DBImproved.cs:11
        public int cf = 0;              // DBImproved.cs:13 -- may be pre-seeded by the caller (FrmMain.cs:1509)
        private IntPtr ctx;

        // ONE native context for the whole application: StartCode creates a DBImproved per work item (FrmMain.cs:2785), and a context
        // owns a CUDA stream and device workspaces that should not be re-created per cell.  The context serialises its calls with an
        // internal mutex, so the pool threads may share it.  Devices = null: GPU 0.  Devices = {0, 1, .., 7}: one context drives all
        // those GPUs -- dbscan() / go_hell_ICP() on large clouds are then spread over them inside libvpc (include/vpc.h, vpc_create).
        private static readonly object gate = new object();
        private static IntPtr shared = IntPtr.Zero;
        public static int[] Devices = null;

        public DBImprovedGpu()
        {
            lock (gate)
            {
                if (shared == IntPtr.Zero)
                    NativeMethods.Check(IntPtr.Zero, NativeMethods.vpc_create(out shared, Devices, Devices == null ? 0 : Devices.Length));
                ctx = shared;
            }
        }
        internal static IntPtr SharedContext() { return new DBImprovedGpu().ctx; }     // ICPGpu and ToolsGpu use the same native context
        public static void Shutdown() { lock (gate) { if (shared != IntPtr.Zero) { NativeMethods.vpc_destroy(shared); shared = IntPtr.Zero; } } }

        public void dbscan(List<Point3D> lst, double e, int minPts)
        {
            int n = lst.Count;
            double[] mx = new double[n], my = new double[n];
            for (int i = 0; i < n; i++) { mx[i] = lst[i].motor_x; my[i] = lst[i].motor_y; }   // getDisP reads only these (DBImproved.cs:16-17)
            int[] cid = new int[n]; byte[] key = new byte[n], cls = new byte[n];
            int amount;
            NativeMethods.Check(ctx, NativeMethods.vpc_dbscan_l1_2d(ctx, mx, my, n, e, minPts, cf, cid, key, cls, out amount));
            for (int i = 0; i < n; i++)
            {
                // The C# never clears flags itself; its callers reset clusterId/isClassed first (FrmMain.cs:1219-1223,
                // 1512-1515; Tools.cs:584-590), so writing all three fields reproduces the post-state exactly.
                lst[i].clusterId = cid[i];
                lst[i].isClassed = cls[i] != 0;
                if (key[i] != 0) lst[i].isKeyPoint = true;   // isKeyPoint is only ever set, never cleared (DBImproved.cs:49)
            }
            pointsAmount += n;        // DBImproved.cs:99
            cf = amount;              // DBImproved.cs:107
            clusterAmount = amount;   // DBImproved.cs:112
        }

        // All cells of the blocked clustering at once: replaces the ThreadPool loop of DoWork3 (FrmMain.cs:1356-1359).
        // Returns clusterAmount per cell; cluster ids written to the points are cell-local like StartCode's.
        public int[] dbscanCells(List<Point3D>[] cells, double e, int minPts)
        {
            long[] off = new long[cells.Length + 1];
            for (int k = 0; k < cells.Length; k++) off[k + 1] = off[k] + (cells[k] == null ? 0 : cells[k].Count);
            int n = (int)off[cells.Length];
            double[] mx = new double[n], my = new double[n];
            for (int k = 0, p = 0; k < cells.Length; k++)
                if (cells[k] != null) foreach (Point3D q in cells[k]) { mx[p] = q.motor_x; my[p] = q.motor_y; p++; }
            int[] cid = new int[n]; byte[] key = new byte[n], cls = new byte[n]; int[] perCell = new int[cells.Length];
            NativeMethods.Check(ctx, NativeMethods.vpc_dbscan_l1_2d_cells(ctx, mx, my, n, off, cells.Length, e, minPts, cid, key, cls, perCell));
            for (int k = 0, p = 0; k < cells.Length; k++)
                if (cells[k] != null) foreach (Point3D q in cells[k])
                {
                    q.clusterId = cid[p]; q.isClassed = cls[p] != 0; if (key[p] != 0) q.isKeyPoint = true; p++;
                }
            return perCell;
        }

        // The whole blocked clustering (MainForm.getClusterFromMotor -> DoWork3 -> CompleteWork3 up to the renumbered labels,
        // FrmMain.cs:1214-1520) in one call.  Writes clusterId / isClassed of every point of rawData, returns MainForm.clusterSum.
        public int dbscanBlocked(List<Point3D> rawData, double e, int minPts, int ptsInCell, out int delSum)
        {
            int n = rawData.Count;
            double[] mx = new double[n], my = new double[n];
            for (int i = 0; i < n; i++) { mx[i] = rawData[i].motor_x; my[i] = rawData[i].motor_y; }
            int[] cid = new int[n];
            int clusterSum, rows, cols; long unassigned;
            NativeMethods.Check(ctx, NativeMethods.vpc_dbscan_blocked_ref(ctx, mx, my, n, e, minPts, ptsInCell, cid, out clusterSum, out delSum,
                                                                          out rows, out cols, out unassigned));
            for (int i = 0; i < n; i++) { rawData[i].clusterId = cid[i]; rawData[i].isClassed = cid[i] != 0; }
            pointsAmount += n;
            clusterAmount = clusterSum;
            return clusterSum;
        }

        // dbscanBlocked plus the list the merge needs: clusForMerge (FrmMain.cs:1517-1520) as indices into rawData with their ids
        public int dbscanBlocked(List<Point3D> rawData, double e, int minPts, int ptsInCell, out int delSum, out long[] mergeOrder, out int[] mergeCid, out long nShared)
        {
            int n = rawData.Count;
            double[] mx = new double[n], my = new double[n];
            for (int i = 0; i < n; i++) { mx[i] = rawData[i].motor_x; my[i] = rawData[i].motor_y; }
            int[] cid = new int[n]; long[] mo = new long[3 * n]; int[] mc = new int[3 * n];
            int clusterSum, rows, cols, sumCells; long unassigned, nMerge;
            NativeMethods.Check(ctx, NativeMethods.vpc_dbscan_blocked_ref_ex(ctx, mx, my, n, e, minPts, ptsInCell, cid, out clusterSum, out delSum, out rows, out cols,
                                                                             out unassigned, out nShared, mo, mc, out nMerge, out sumCells));
            for (int i = 0; i < n; i++) { rawData[i].clusterId = cid[i]; rawData[i].isClassed = cid[i] != 0; }
            mergeOrder = new long[nMerge]; mergeCid = new int[nMerge];
            Array.Copy(mo, mergeOrder, nMerge); Array.Copy(mc, mergeCid, nMerge);
            pointsAmount += n;
            clusterAmount = clusterSum;
            return clusterSum;
        }

        // The centroid merge of Clustering.MergeBtn_Click (Clustering.cs:141-153) on the result of dbscanBlocked: clusForMerge comes back from
        // vpc_dbscan_blocked_ref_ex as point indices; the per-entry arrays are gathered in list order (the order the C# sums centroids in).
        public int mergeByDistance(List<Point3D> rawData, long[] mergeOrder, int[] mergeCid, int clusterSum, double thre)
        {
            int k = mergeOrder.Length;
            double[] xyz = new double[3 * k], mx = new double[k], my = new double[k];
            for (int t = 0; t < k; t++)
            {
                Point3D p = rawData[(int)mergeOrder[t]];
                xyz[t] = p.X; xyz[k + t] = p.Y; xyz[2 * k + t] = p.Z; mx[t] = p.motor_x; my[t] = p.motor_y;
            }
            int[] newCid = new int[k], dfrom = new int[clusterSum], dto = new int[clusterSum], cids = new int[clusterSum];
            double[] c5 = new double[5 * clusterSum], nc5 = new double[5 * clusterSum];
            int newAmount, nDict, nCenters;
            NativeMethods.Check(ctx, NativeMethods.vpc_merge_ids_by_distance(ctx, mergeCid, xyz, mx, my, k, clusterSum, thre, newCid, out newAmount,
                                                                             dfrom, dto, out nDict, c5, cids, out nCenters, nc5));
            for (int t = 0; t < k; t++) rawData[(int)mergeOrder[t]].clusterId = newCid[t];      // Tools.cs:557-560
            return newAmount;
        }

        public void Dispose() { ctx = IntPtr.Zero; }   // the native context is shared: DBImprovedGpu.Shutdown() releases it at application exit
    }
}
